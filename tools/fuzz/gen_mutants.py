import numpy as np, oracle, os, io
from PIL import Image
rng=np.random.default_rng(11)
os.makedirs('/tmp/fz/in',exist_ok=True)
b=io.BytesIO(); Image.fromarray(oracle.synth_image(70,50,3)).save(b,"JPEG",subsampling=2,quality=60,optimize=True)
seeds=[oracle.oracle_encode(oracle.synth_image(64,48,3),1,75,1,restart=4), oracle.oracle_encode(oracle.synth_image(40,40,3),0,3,0), oracle.oracle_encode(oracle.synth_image(50,30,1),1,85,0,restart=24), open('/root/repo/tests/golden/data_test.jpg','rb').read(), b.getvalue()]
for it in range(4000):
    s=bytearray(seeds[it%len(seeds)])
    for _ in range(rng.integers(1,6)):
        op=rng.integers(0,5)
        if op==0: s[rng.integers(0,len(s))]=rng.integers(0,256)
        elif op==1 and len(s)>10:
            i=rng.integers(0,len(s)); del s[i:i+rng.integers(1,40)]
        elif op==2: i=rng.integers(0,len(s)); s[i:i]=bytes(rng.integers(0,256,rng.integers(1,8),dtype=np.uint8))
        elif op==3: i=rng.integers(0,max(1,len(s)-2)); s[i]=0xFF; s[i+1]=rng.choice([0xD0,0xD9,0xC4,0xDA,0xDD,0x00,0xC0,0xDB])
        elif len(s) > 3: s=s[:rng.integers(2,len(s))]
    open('/tmp/fz/in/%04d.jpg'%it,'wb').write(bytes(s))
print("generated")
