"""oracle/ -- CPU checkers for the JPEG-encode path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  imagecodecs_b200 (the product) never does.
"""
from .oracle import (  # noqa: F401
    QMODE_TJE, QMODE_IJG, SUB_444, SUB_420,
    build, have_ref, oracle_encode, oracle_stages, oracle_headers, oracle_tables,
    ref_encode, ref_decode, read_bmp, num_blocks, restart_interval,
)
from .synth import synth_image, synth_batch  # noqa: F401
