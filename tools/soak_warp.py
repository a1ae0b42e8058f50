#!/usr/bin/env python
"""Randomised soak of the arbitrary-angle loader's arithmetic on the CPU (test infrastructure): builds
tests/warp_host.cpp (the header loader_affine_kernel is built from) and compares it with cv2.warpAffine on random
rotations, similarity transforms and arbitrary 2x3 matrices, with flips and crops.

    python tools/soak_warp.py SEED SECONDS        # round 1: 246k cases over two seeds, 0 mismatches
"""
import ctypes, os, subprocess, sys, tempfile, time
import numpy as np, cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_np as O
so = os.path.join(tempfile.mkdtemp(), "libwarp_host.so")
subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", "-I",
                os.path.join(ROOT, "recursion_cellular_image_classification_b200", "csrc"),
                os.path.join(ROOT, "tests", "warp_host.cpp"), "-o", so], check=True)
lib = ctypes.CDLL(so)
rng=np.random.default_rng(int(sys.argv[1])); T=float(sys.argv[2])
bad=0;n=0;t0=time.time()
while time.time()-t0<T:
    H,W=int(rng.integers(1,120)),int(rng.integers(1,120))
    src=rng.integers(0,256,size=(6,H,W),dtype=np.uint8)
    mode=rng.integers(3)
    if mode==0: M=O.rotation_matrix(W,H,float(rng.uniform(-180,180)))
    elif mode==1:
        M=O.rotation_matrix(W,H,float(rng.uniform(-180,180)),scale=float(rng.uniform(0.2,5)))
        M[:,2]+=rng.uniform(-300,300,size=2)
    else:
        M=rng.uniform(-2,2,size=(2,3)); M[:,2]=rng.uniform(-W,W),rng.uniform(-H,H)
        if abs(np.linalg.det(M[:,:2]))<0.05: continue
    vflip,hflip=int(rng.integers(2)),int(rng.integers(2))
    Ho,Wo=int(rng.integers(1,H+1)),int(rng.integers(1,W+1)); y0,x0=int(rng.integers(0,H-Ho+1)),int(rng.integers(0,W-Wo+1))
    dst=np.empty((6,Ho,Wo),np.uint8)
    lib.warp_host_planar_u8(src.ctypes.data_as(ctypes.c_void_p),H,W,np.ascontiguousarray(M).ctypes.data_as(ctypes.c_void_p),vflip,hflip,y0,x0,Ho,Wo,dst.ctypes.data_as(ctypes.c_void_p))
    img=np.moveaxis(src,0,2)
    if vflip: img=img[::-1]
    if hflip: img=img[:,::-1]
    ref=cv2.warpAffine(np.ascontiguousarray(img),M,(W,H),flags=cv2.INTER_LINEAR,borderMode=cv2.BORDER_REFLECT_101)
    if ref.ndim==2: ref=ref[...,None]
    ref=ref[y0:y0+Ho,x0:x0+Wo]
    if not np.array_equal(np.moveaxis(dst,0,2),ref):
        bad+=1; print("BAD",H,W,mode,M.tolist())
    n+=1
print("cases",n,"bad",bad)
