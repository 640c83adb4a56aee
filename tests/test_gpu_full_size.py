"""GPU (-m gpu): BASELINE.json's configurations at their full sizes, every image (or a stratified sample of a real
full-size launch) compared byte for byte with the CPU checker; the fused round-1 kernel against the split pipeline; two
host threads inside the C ABI at once; the single-process multi-GPU path where the box has more than one GPU.

The CPU side runs on a thread pool (ctypes releases the GIL): oracle/_ref -- the compiled, unmodified jpeg_enc.h -- for the
native modes, the oracle port for the extended ones."""
import concurrent.futures as cf
import os
import subprocess
import sys
import threading

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libtje_ref.so"))


def cpu_encode(img, qm, q, sub):
    if qm == 0 and sub == 0 and img.shape[2] in (3, 4) and HAVE_REF:
        rc, ref = oracle.ref_encode(img, q)
        assert rc == 1
        return ref
    return oracle.oracle_encode(img, qm, q, sub)


def compare_all(files, host, modes, indices):
    """files[i] == CPU checker on host[i] for every i of indices (threaded)."""
    with cf.ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 4)) as ex:
        want = list(ex.map(lambda i: cpu_encode(host[i], *modes(i)), indices))
    bad = [i for i, w in zip(indices, want) if files[i] != w]
    assert not bad, "images differ from the CPU checker: %s" % bad[:10]


def test_config2_every_image_of_the_256_batch(gpu):
    """BASELINE configs[1]: 256 x 1920x1080 RGB, IJG q75 4:2:0 -- the bench workload -- and its byte-pinned native twin
    (tje quality 2, 4:4:4, checked against the compiled reference): EVERY image of the batch."""
    from imagecodecs_b200.synth import synth_batch
    dev = synth_batch(256, 1920, 1080, 3, device="cuda")
    torch.cuda.synchronize()
    host = dev.cpu().numpy()
    imgs = [dev[i] for i in range(256)]
    for (qm, q, sub) in [(1, 75, 1), (0, 2, 0)]:
        plan = gpu.Plan.for_arrays(imgs, qm, q, sub, device=0)
        plan.run()
        files = plan.fetch()
        plan.close()
        compare_all(files, host, lambda i: (qm, q, sub), range(256))


def test_config3_every_image_of_a_4k_shard(gpu):
    """BASELINE configs[2]: 3840x2160 RGB, q=90 4:4:4; the 16-image shard one of eight GPUs gets (few images in the
    launch: the deferred write-out variant of the entropy kernel), and the native twin tje 3."""
    from imagecodecs_b200.synth import synth_batch
    dev = synth_batch(16, 3840, 2160, 3, first=112, device="cuda")
    torch.cuda.synchronize()
    host = dev.cpu().numpy()
    for (qm, q, sub) in [(1, 90, 0), (0, 3, 0)]:
        plan = gpu.Plan.for_arrays([dev[i] for i in range(16)], qm, q, sub, device=0)
        plan.run()
        files = plan.fetch()
        plan.close()
        compare_all(files, host, lambda i: (qm, q, sub), range(16))


def test_config5_real_16384_image_launch_stratified_sample(gpu):
    """BASELINE configs[4]: 16384 x 512x512 RGB, quality by n % 3 -> {50, 75, 95}, 4:2:0, in ONE plan (three quantiser groups,
    each a launch of 5461 / 5462 images): 384 images spread over the whole batch (every 43rd, which walks through all
    three qualities, plus the first and last 16) against the oracle, and every status / size sane."""
    from imagecodecs_b200.synth import synth_batch
    n = 16384
    dev = synth_batch(n, 512, 512, 3, device="cuda", chunk=1024)
    torch.cuda.synchronize()
    q = [[50, 75, 95][i % 3] for i in range(n)]
    plan = gpu.Plan.for_arrays([dev[i] for i in range(n)], 1, q, 1, device=0)
    plan.run()
    sizes = [plan.encoded_size(i) for i in range(n)]
    assert min(sizes) > 700 and max(sizes) < 512 * 512 * 3
    files = plan.fetch()
    plan.close()
    sample = sorted(set(list(range(0, n, 43)) + list(range(16)) + list(range(n - 16, n))))
    assert len(sample) >= 384 and {i % 3 for i in sample} == {0, 1, 2}
    host = {i: dev[i].cpu().numpy() for i in sample}
    compare_all(files, host, lambda i: (1, q[i], 1), sample)
    # the native twin of the same launch shape on a 2048-image slice: tje {1, 2, 3} by n % 3, 4:4:4
    tq = [1 + i % 3 for i in range(2048)]
    twin = gpu.Plan.for_arrays([dev[i] for i in range(2048)], 0, tq, 0, device=0)
    twin.run()
    tfiles = twin.fetch()
    twin.close()
    tsample = list(range(0, 2048, 31))
    compare_all(tfiles, {i: dev[i].cpu().numpy() for i in tsample}, lambda i: (0, tq[i], 0), tsample)


def test_fused_round1_kernel_and_split_pipeline_write_the_same_bytes(gpu, monkeypatch):
    """JPEG_GPU_PIPELINE=fused selects the single pass-1 kernel of round 1 (kept for the A/B numbers): all layouts, restart
    intervals and a slow-path image give the same files through both, and both equal the oracle."""
    rng = np.random.default_rng(5)
    imgs = [oracle.synth_image(200, 120, 3, n=1), oracle.synth_image(64, 64, 1), oracle.synth_image(333, 201, 4, n=2),
            rng.integers(0, 256, size=(96, 160, 3), dtype=np.uint8), oracle.synth_image(640, 360, 3, n=3)]
    qm, q, sub = [0, 1, 1, 0, 1], [2, 85, 75, 3, 50], [0, 0, 1, 0, 1]
    res = {}
    for mode in ("split", "fused"):
        if mode == "fused": monkeypatch.setenv("JPEG_GPU_PIPELINE", "fused")
        else: monkeypatch.delenv("JPEG_GPU_PIPELINE", raising=False)
        plan = gpu.Plan.for_arrays([torch.from_numpy(i).cuda() for i in imgs], qm, q, sub, device=0)
        assert plan.fused == (mode == "fused")
        plan.run()
        res[mode] = plan.fetch()
        plan.close()
        res[mode + "_restart"] = gpu.encode_batch(imgs, qm, q, sub, device=0, flags=gpu.FLAG_RESTART)[0]
    monkeypatch.delenv("JPEG_GPU_PIPELINE", raising=False)
    assert res["split"] == res["fused"] and res["split_restart"] == res["fused_restart"]
    for i in range(len(imgs)):
        assert res["split"][i] == oracle.oracle_encode(imgs[i], qm[i], q[i], sub[i])


def test_two_host_threads_inside_the_library_at_once(gpu):
    """The reference encoder is re-entrant (all state in a stack TJEState, jpeg_enc.h:1228): two threads that call
    jpeg_gpu_encode_batch concurrently (they share the device's stream ring and the block pool) both get their own bytes."""
    batches = [[oracle.synth_image(640, 360, 3, n=10 * t + k) for k in range(12)] for t in range(2)]
    modes = [(1, 75, 1), (0, 2, 0)]
    out, err = [None, None], []

    def work(t):
        try:
            for _ in range(4):
                files, st = gpu.encode_batch(batches[t], *modes[t], device=0)
                assert st == [0] * 12
                out[t] = files
        except Exception as e:      # surfaced below
            err.append(e)

    ths = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for th in ths: th.start()
    for th in ths: th.join()
    assert not err, err
    for t in range(2):
        for k in range(12):
            assert out[t][k] == oracle.oracle_encode(batches[t][k], *modes[t]), (t, k)


def test_contiguous_host_batches_upload_as_one_copy_with_the_same_files(gpu):
    """jpeg_gpu_encode_batch uploads images that follow one another in host memory (and in the plan's pixel arena) with ONE
    copy per run: slices of one array, the same images as separate arrays, a run broken by an image of another size, by a
    bottom-up image and by an image whose size is not a multiple of the arena's alignment all give the oracle's bytes."""
    rng = np.random.default_rng(11)
    big = np.stack([oracle.synth_image(256, 128, 3, n=i) for i in range(6)])           # 98304 bytes each: a multiple of 256
    odd = np.stack([oracle.synth_image(250, 125, 3, n=20 + i) for i in range(4)])       # 93750 bytes: runs break at every image
    other = oracle.synth_image(96, 64, 3, n=40)
    batch = [big[0], big[1], big[2], other, big[3], big[4], big[5], odd[0], odd[1], odd[2], odd[3]]
    want = [oracle.oracle_encode(im, 1, 85, 1) for im in batch]
    files, st = gpu.encode_batch(batch, 1, 85, 1, device=0)                              # slices: pointers are contiguous
    assert st == [0] * len(batch) and files == want
    files, st = gpu.encode_batch([im.copy() for im in batch], 1, 85, 1, device=0)        # separate allocations
    assert st == [0] * len(batch) and files == want
    # bottom-up rows: the same memory read from the last row upwards == the flipped image
    files, st = gpu.encode_batch([big[i] for i in range(6)], 0, 2, 0, device=0, bottom_up=True)
    assert st == [0] * 6 and files == [oracle.oracle_encode(np.ascontiguousarray(big[i][::-1]), 0, 2, 0) for i in range(6)]
    del rng


MULTI_GPU_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import oracle, imagecodecs_b200 as jg
G = jg.init(None)
assert G >= 2, G
n = 8 * G + 3
imgs = [oracle.synth_image(512, 512, 3, n=i) for i in range(n)]
q = [[50, 75, 95][i %% 3] for i in range(n)]
whole, st = jg.encode_batch(imgs, 1, q, 1, device=-1)          # one host thread + stream ring per GPU, shards by image index
assert st == [0] * n
one, st1 = jg.encode_batch(imgs, 1, q, 1, device=G - 1)         # the same batch on the last GPU alone
assert one == whole
for i in range(0, n, 3):
    assert whole[i] == oracle.oracle_encode(imgs[i], 1, q[i], 1), i
big = [oracle.synth_image(3840, 2160, 3, n=i) for i in range(G)]
files, st = jg.encode_batch(big, 0, 3, 0, device=-1)
assert st == [0] * G and files[G - 1] == jg.encode_batch(big[G - 1:], 0, 3, 0, device=0)[0][0]
print("multi-gpu ok", G)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="one GPU on this box")
def test_single_process_multi_gpu_batch(gpu):
    """jpeg_gpu_encode_batch(device = -1) on every GPU of the box (contiguous index ranges, no collective): the
    files do not depend on how many GPUs shared the batch.  Own process: the suite's library handle is bound to GPU 0."""
    r = subprocess.run([sys.executable, "-c", MULTI_GPU_SCRIPT % ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "multi-gpu ok" in r.stdout, r.stdout + r.stderr
