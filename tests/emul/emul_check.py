import sys, ctypes, struct, time; sys.path.insert(0,'/root/repo')
from hvqm4_b200 import synth
from oracle.bindings import RefDecoder
import numpy as np
lib = ctypes.CDLL('/root/repo/tests/emul/libhvqm4_emul.so')
lib.h4e_seq_create.restype = ctypes.c_void_p
lib.h4e_seq_create.argtypes = [ctypes.c_int]*5
lib.h4e_parse_begin.restype = ctypes.c_size_t
lib.h4e_parse_begin.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
lib.h4e_parse_finish.restype = ctypes.c_uint32
lib.h4e_parse_finish.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
lib.h4e_seq_destroy.argtypes = [ctypes.c_void_p]
lib.emul_recon_picture.argtypes = [ctypes.c_void_p]*4

def demux(data):
    ver = 15 if data[:9]==b'HVQM4 1.5' else 13
    ngop = struct.unpack('>I', data[0x18:0x1c])[0]
    w,h = struct.unpack('>HH', data[0x34:0x38])
    pos = 0x44; recs=[]
    for g in range(ngop):
        nv, na = struct.unpack('>II', data[pos+8:pos+16]); pos += 20
        while nv or na:
            id1,id2,size = struct.unpack('>HHI', data[pos:pos+8]); pos += 8
            if id1==1: recs.append((id2, data[pos:pos+size])); nv-=1
            else: na-=1
            pos += size
    return ver,w,h,recs

def run(w,h,v,gop,prof,seed,ngop=2):
    d = synth.generate(w,h,v,gop,ngop,seed=seed,profile=prof)
    ver,W,H,recs = demux(d)
    seq = lib.h4e_seq_create(W,H,2,2,int(ver==15))
    fb = W*H*3//2
    bufs = [np.zeros(fb+64,np.uint8) for _ in range(3)]
    past,present,future = 0,1,2
    ref = RefDecoder(d)
    ok=True; tparse=0; blobsz=0
    for i,(ty,rec) in enumerate(recs):
        if ty!=0x30: past,future = future,past
        pic = rec[4:]+b'\0'*8
        t0=time.perf_counter()
        n = lib.h4e_parse_begin(seq, ty, pic, len(pic))
        blob = np.zeros(n,np.uint8)
        err = lib.h4e_parse_finish(seq, blob.ctypes.data)
        tparse += time.perf_counter()-t0
        blobsz += n
        fut = bufs[present] if ty==0x20 else bufs[future]
        rc = lib.emul_recon_picture(blob.ctypes.data, bufs[present].ctypes.data, bufs[past].ctypes.data, fut.ctypes.data)
        rt,_,_,yuv = ref.decode_next()
        got = bufs[present][:fb].tobytes()
        if rc or err or got!=yuv:
            ok=False
            x=np.frombuffer(yuv,np.uint8); y=bufs[present][:fb]
            idx=np.nonzero(x!=y)[0]
            print('  frame',i,hex(ty),'rc',rc,'err',err,'ndiff',len(idx), 'first', idx[:8])
            break
        if ty!=0x30: present,future = future,present
    lib.h4e_seq_destroy(seq)
    print(w,h,v,gop,prof,'EMUL==REF' if ok else 'MISMATCH', 'parse fps %.0f'%(len(recs)/tparse), 'blob KB/frame %.1f'%(blobsz/len(recs)/1024))
    return ok

if __name__=='__main__':
    allok=True
    for args in [(320,240,15,"IIII",0,1),(320,240,15,"IPBBPBB",0,1),(640,480,15,"IPPPP",0,2),(640,480,13,"IPBBPBB",1,3),(320,240,13,"IPBB",0,4),(280,152,15,"IPB",0,5),(1024,768,13,"IPB",0,6),(640,480,15,"I"+"PBB"*5,1,7),(648,488,15,"IPB",0,8)]:
        allok &= run(*args)
    print('ALL OK' if allok else 'FAILURES')
