import sys, torch
sys.path.insert(0, "/root/repo")
import bench
from tpp_b200 import _lib
from tpp_b200.common.storage import Storage
for T, N in ((256, 4096), (256, 256), (256, 8192), (256, 16384)):
    st = Storage((1,), 1, T, N, "cuda")
    st.rew.normal_(); st.value.normal_()
    s = _lib.stream_ptr
    f_gae = lambda: _lib.call("tpp_gae", _lib.ptr(st.rew), _lib.ptr(st.done_u8), _lib.ptr(st.value), _lib.ptr(st.adv), _lib.ptr(st.ret), _lib.ptr(st.moments), T, N, st.ld, 0.99, 0.95, s())
    f_norm = lambda: _lib.call("tpp_adv_normalize", _lib.ptr(st.adv), _lib.ptr(st.moments), T, N, st.ld, s())
    f_zero = lambda: st.moments.zero_()
    print(T, N, "gae %.2f us  normalize %.2f us  zero %.2f us  all %.2f us" % (
        bench.time_kernel(f_gae, 20) * 1e6, bench.time_kernel(f_norm, 20) * 1e6, bench.time_kernel(f_zero, 20) * 1e6,
        bench.time_kernel(lambda: st.compute_estimates(0.99, 0.95, True, True), 20) * 1e6))
