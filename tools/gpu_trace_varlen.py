"""Timeline of the varlen attention kernel's hand-over points (debug library built with -DEMDR2_VARLEN_TRACE, see
csrc/attention_varlen.cu): CTA 0's first softmax warp and MMA thread, per key block, in SM clocks."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from emdr2_b200 import _lib, ops
from emdr2_b200.packed import PackedBatch

DEV = torch.device("cuda:0")
_lib.load()
trace = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ab", "libemdr2_trace.so"))
for name, (restype, argtypes) in _lib._SIGNATURES.items():
    fn = getattr(trace, name)
    fn.restype, fn.argtypes = restype, argtypes
_lib._LIB = trace
heads, h = 12, 768
lens = np.random.RandomState(0).randint(380, 513, size=400)
pb = PackedBatch(lens, 512, heads, DEV)
qkv = (torch.randn(pb.T, 3 * h, device=DEV) * 0.5).to(torch.bfloat16)
o = torch.empty(pb.T, h, dtype=torch.bfloat16, device=DEV)
for _ in range(3):
    ops.attention_varlen(qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:], heads, pb.items, pb.n_items, scale=0.125, out=o)
torch.cuda.synchronize()
buf = np.zeros((2, 64, 8), dtype=np.int64)
rc = trace.emdr2_varlen_trace_read(buf.ctypes.data_as(ctypes.c_void_p), buf.size)
assert rc == 0, rc
sm, mm = buf[0], buf[1]
t0 = sm[0, 0]
print("softmax warp (CTA 0): per key block, clocks relative to the first stamp")
print("blk  wait_S   S_ready  ld_done  max_done pv_ok    exp+st   | block period | (item end: o_wait, o_ready)")
for b in range(40):
    r = sm[b] - t0
    per = sm[b + 1, 0] - sm[b, 0]
    print("%3d  %7d  %7d  %7d  %7d  %7d  %7d | %6d | %s" % (b, r[0], r[1], r[2], r[3], r[4], r[5], per,
          ("%d %d" % (r[6], r[7])) if sm[b, 6] else ""))
d = np.diff(sm[4:40, :6], axis=1)
print("mean segment clocks (blocks 4..39): wait S %.0f, tcgen05.ld %.0f, max %.0f, wait pv_done %.0f, exp + P store %.0f; block period %.0f"
      % (d[:, 0].mean(), d[:, 1].mean(), d[:, 2].mean(), d[:, 3].mean(), d[:, 4].mean(), np.diff(sm[4:40, 0]).mean()))
print("MMA thread: per block g: loop top, s_free seen, S(g+1) issued, p_full seen, P.V issued (relative)")
for g in range(12):
    print(g, (mm[g, :5] - t0).tolist())
dm = np.diff(mm[4:40, :5], axis=1)
print("mean MMA segments: wait s_free %.0f, issue S %.0f, wait p_full %.0f, issue PV %.0f" % tuple(dm.mean(0)))
