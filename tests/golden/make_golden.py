"""Regenerate tests/golden/* from the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference/jpeg_enc.h + jpeg_dec.h).  Run in the dev container only:

    python tests/golden/make_golden.py

Outputs
  kat.json            known answers (size + SHA-256 of the complete JPEG the reference writes)
  ref_*.jpg           the reference's own output bytes for a few tiny inputs
  fixture_pixels.npz  pixels of the reference's data/ fixtures as the codecs.h path sees them:
                      cat_bgr  = data/cat.bmp through readBmp semantics (codecs.cpp:255-320, B,G,R kept)
                      testjpg  = data/test.jpg through njDecode (what tests.cpp pass 1 feeds writeJpg)
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

REF = "/root/reference/data"


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    cat = oracle.read_bmp(open(os.path.join(REF, "cat.bmp"), "rb").read())
    testbmp = oracle.read_bmp(open(os.path.join(REF, "test.bmp"), "rb").read())
    testjpg = oracle.ref_decode(open(os.path.join(REF, "test.jpg"), "rb").read())
    np.savez_compressed(os.path.join(HERE, "fixture_pixels.npz"), cat_bgr=cat, testjpg=testjpg, testbmp_bgr=testbmp)
    # the decoder's input fixture: the reference's own data/test.jpg (13 KB, 4:2:0), whose njDecode output is `testjpg` above
    import shutil
    shutil.copyfile(os.path.join(REF, "test.jpg"), os.path.join(HERE, "data_test.jpg"))

    kat = []

    def add(name, img, q, gen=None, save=False):
        rc, data = oracle.ref_encode(img, q)
        assert rc == 1
        e = {"name": name, "w": int(img.shape[1]), "h": int(img.shape[0]), "ncomp": int(img.shape[2]),
             "tje_quality": q, "bytes": len(data), "sha256": sha(data), "gen": gen}
        if save:
            fn = "ref_%s_q%d.jpg" % (name, q)
            open(os.path.join(HERE, fn), "wb").write(data)
            e["file"] = fn
        kat.append(e)

    for q in (1, 2, 3):
        add("cat_bgr", cat, q, {"fixture": "cat_bgr"})
    add("testjpg", testjpg, 3, {"fixture": "testjpg"})
    add("testbmp_bgr", testbmp, 3, {"fixture": "testbmp_bgr"})
    S = oracle.synth_image
    g = lambda w, h, c, n, kind: {"synth": [w, h, c, n, kind]}
    add("photo_8x8", S(8, 8, 3, 0, "photo"), 3, g(8, 8, 3, 0, "photo"), save=True)
    for q in (1, 2, 3):
        add("photo_17x13", S(17, 13, 3, 0, "photo"), q, g(17, 13, 3, 0, "photo"), save=True)
    add("photo_17x13_rgba", S(17, 13, 4, 0, "photo"), 3, g(17, 13, 4, 0, "photo"), save=True)
    add("noise_64x64", S(64, 64, 3, 0, "noise"), 3, g(64, 64, 3, 0, "noise"), save=True)
    add("noise_64x64", S(64, 64, 3, 0, "noise"), 1, g(64, 64, 3, 0, "noise"), save=True)
    add("photo_512_n5", S(512, 512, 3, 5, "photo"), 2, g(512, 512, 3, 5, "photo"))
    for q in (1, 2, 3):
        add("photo_1080p", S(1920, 1080, 3, 0, "photo"), q, g(1920, 1080, 3, 0, "photo"))
    add("noise_1080p", S(1920, 1080, 3, 0, "noise"), 3, g(1920, 1080, 3, 0, "noise"))
    add("photo_1921x1083", S(1921, 1083, 3, 0, "photo"), 2, g(1921, 1083, 3, 0, "photo"))
    add("photo_4k", S(3840, 2160, 3, 0, "photo"), 3, g(3840, 2160, 3, 0, "photo"))
    json.dump({"source": "oracle/_ref/libtje_ref.so = /root/reference/jpeg_enc.h, gcc -std=gnu99 -O2 -ffp-contract=off",
               "entries": kat}, open(os.path.join(HERE, "kat.json"), "w"), indent=1)
    print("wrote %d known answers" % len(kat))


if __name__ == "__main__":
    main()
