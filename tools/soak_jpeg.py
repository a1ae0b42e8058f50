#!/usr/bin/env python
"""Randomised soak of the device JPEG decoder's arithmetic on the CPU (test infrastructure): builds tests/jpeg_host.cpp
(the header the kernels are built from, single-lane flow and a lane-by-lane emulation of the all-lanes flow) and
compares both with cv2.imdecode on random images / qualities / table and restart options.

    python tools/soak_jpeg.py SEED SECONDS        # round 1: 269k cases over three seeds, 0 mismatches
"""
import ctypes, os, subprocess, sys, tempfile, time
import numpy as np, cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(tempfile.mkdtemp(), "libjpeg_host.so")
subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-I",
                os.path.join(ROOT, "recursion_cellular_image_classification_b200", "csrc"),
                os.path.join(ROOT, "tests", "jpeg_host.cpp"), "-o", so], check=True)
lib = ctypes.CDLL(so)
def dec(buf,H,W,par):
    a=np.frombuffer(buf,dtype=np.uint8); dst=np.zeros((H,W),np.uint8); r=ctypes.c_int(0)
    if par: st=lib.jpeg_host_decode_gray_parallel(a.ctypes.data_as(ctypes.c_void_p),len(a),H,W,dst.ctypes.data_as(ctypes.c_void_p),ctypes.byref(r))
    else: st=lib.jpeg_host_decode_gray(a.ctypes.data_as(ctypes.c_void_p),len(a),H,W,dst.ctypes.data_as(ctypes.c_void_p))
    return st,dst
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
bad=0; n=0; t0=time.time()
while time.time()-t0 < float(sys.argv[2]) if len(sys.argv)>2 else 60:
    H,W=int(rng.integers(1,200)),int(rng.integers(1,200))
    kind=rng.integers(5)
    if kind==0: img=rng.integers(0,256,size=(H,W),dtype=np.uint8)
    elif kind==1: img=np.clip(rng.gamma(2.0,rng.uniform(1,40),size=(H,W)),0,255).astype(np.uint8)
    elif kind==2: img=cv2.GaussianBlur(rng.integers(0,256,size=(H,W),dtype=np.uint8),(0,0),float(rng.uniform(0.5,6)))
    elif kind==3: img=np.full((H,W),int(rng.integers(256)),np.uint8); img[rng.integers(H):,rng.integers(W):]=int(rng.integers(256))
    else:
        img=(rng.random((H,W))<rng.uniform(0.001,0.2)).astype(np.uint8)*int(rng.integers(1,256))
    q=int(rng.integers(1,101)); params=[cv2.IMWRITE_JPEG_QUALITY,q]
    if rng.random()<0.3: params+=[cv2.IMWRITE_JPEG_OPTIMIZE,1]
    if rng.random()<0.2: params+=[cv2.IMWRITE_JPEG_RST_INTERVAL,int(rng.integers(1,40))]
    ok,buf=cv2.imencode(".jpg",img,params); ref=cv2.imdecode(buf,-1); b=buf.tobytes()
    for par in (False,True):
        st,got=dec(b,H,W,par)
        if par and st==-1: continue
        if st!=0 or not np.array_equal(got,ref):
            bad+=1; print("BAD",H,W,kind,q,params,par,st); np.save("bad_%d.npy"%n,img)
    n+=1
print("cases",n,"bad",bad)
