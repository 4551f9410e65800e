"""Host-side index packing (csrc/trmf_b200.cu: pack_bitmap_slabs) timed on its own: three algorithms x worker counts."""
import ctypes, numpy as np, time, os, sys
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'exp-trmf-nips16_b200', 'trmf', 'corelib', 'trmf_float64.so'))
fn = lib.trmf_b200_pack_bitmap_host
fn.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
T, n = 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 3000
rng = np.random.RandomState(0)
mask = rng.rand(n, T) < 0.9
col_ptr = np.r_[0, np.cumsum(mask.sum(1))].astype(np.uint64)
row_idx = np.nonzero(mask)[1].astype(np.uint32)
out = np.zeros(n*((T+31)//32), dtype=np.uint32)
ref=None
import numpy as np
for thr in (2, 9, 16, 17):
    os.environ['TRMF_B200_PACK_THREADS'] = str(thr)
    for algo in ('words','or8'):
        os.environ['TRMF_B200_PACK_ALGO']=algo
        ts=[]
        for r in range(5):
            out[:]=0xdeadbeef
            t0=time.perf_counter(); rc=fn(T,n,col_ptr.ctypes.data,row_idx.ctypes.data,out.ctypes.data); ts.append(time.perf_counter()-t0)
        if ref is None: ref=out.copy()
        print(f"threads {thr} (workers {thr-1}) {algo}: {min(ts)*1e3:.2f} ms = {min(ts)*1e9/len(row_idx)*(thr-1):.3f} ns/entry/worker, rc {rc}, same {np.array_equal(ref,out)}")
