"""Measured peaks for the roofline of the fp64 kernels: the DFMA rate of the box and the
shared-memory wavefront cost of 128-bit loads, through the probe kernels of the library (csrc/probes.cu).
Measurement utilities; nothing on the product path calls them."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def fp64_peak(device, iters=4096, ctas_per_sm=8, repeats=3):
    """{'tflops': from CUDA events over the whole launch, 'dfma_per_clk_per_sm': from the CTAs' own clock64 spans,
    'sm_count', 'ms'}: 256-thread CTAs, 16 independent DFMA chains per thread (the pipe is saturated at 8 CTAs per SM)."""
    dev = torch.device(device)
    with torch.cuda.device(dev):
        lib = _lib.lib()
        sm, _, _ = _lib.device_info()
        ctas = sm * ctas_per_sm
        seed = torch.tensor([1.0000001, 1.0e-9], dtype=torch.float64, device=dev)
        sink = torch.zeros(1, dtype=torch.float64, device=dev)
        cyc = torch.zeros(ctas, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream(dev)
        best = None
        for _ in range(repeats + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            n = lib.dn_probe_fp64(ctas, iters, C.c_void_p(seed.data_ptr()), C.c_void_p(sink.data_ptr()),
                                  C.c_void_p(cyc.data_ptr()), C.c_void_p(st.cuda_stream))
            e1.record(st)
            torch.cuda.synchronize(dev)
            if n < 0:
                _lib.check(int(n))
            ms = e0.elapsed_time(e1)
            if best is None or ms < best:
                best = ms
        tflops = 2.0 * n / (best * 1e-3) / 1e12
        # per clock and SM at the SM clock the device reports as its maximum (the probe runs for milliseconds at boost)
        mhz = torch.cuda.get_device_properties(dev).clock_rate / 1000.0 if hasattr(torch.cuda.get_device_properties(dev), "clock_rate") else 1965.0
        per_clk = n / (best * 1e-3) / sm / (mhz * 1e6)
        return dict(tflops=tflops, dfma_per_clk_per_sm=per_clk, sm_mhz_assumed=mhz, sm_count=sm, ms=best)


def lds_wavefronts(device, patterns=range(14), iters=2000):
    """{pattern: SM cycles per warp-level 128-bit shared load} with 8 warps of one CTA per SM issuing back to back
    (= shared-memory wavefronts per request once the pipe is the bottleneck)."""
    dev = torch.device(device)
    out = {}
    with torch.cuda.device(dev):
        lib = _lib.lib()
        sm, _, _ = _lib.device_info()
        sink = torch.zeros(1, dtype=torch.float64, device=dev)
        cyc = torch.zeros(sm, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream(dev)
        for pat in patterns:
            n = lib.dn_probe_lds(sm, int(pat), iters, C.c_void_p(sink.data_ptr()), C.c_void_p(cyc.data_ptr()),
                                 C.c_void_p(st.cuda_stream))
            torch.cuda.synchronize(dev)
            if n < 0:
                _lib.check(int(n))
            out[int(pat)] = float(np.median(cyc.cpu().numpy())) / float(n)
    return out
