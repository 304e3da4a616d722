"""Volumetric fog accumulation on the GPU (granite_b200/csrc/grb_fog.cu through the C ABI) against the oracle.  Sorted after the
validated tests and expected-to-fail-tolerant: written after the round's GPU time had run out.  Verified without a GPU: the
kernel's source compiled for the CPU, bit for bit with the oracle, and the oracle against the reference's shader
(tests/test_fog_cpu.py).  On hardware exp2f is CUDA's: the stored fp16 values may differ by one ulp.  An XPASS means the first
hardware run agreed."""
import ctypes as C

import numpy as np
import pytest

from tests.test_fog_cpu import make_density

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="first run on hardware: the kernel is verified through CPU emulation of its source only")]


@pytest.mark.parametrize("w,h,d", [(33, 17, 7), (160, 92, 64), (320, 180, 128)])
def test_cuda_fog_accumulate_vs_oracle(cuda, oracle, w, h, d):
    import torch

    from granite_b200 import capi, harness

    light = make_density(w, h, d, seed=2)
    dev = harness.to_dev(light)
    fog = torch.zeros((d, h, w, 4), dtype=torch.int16, device="cuda")
    capi.check(capi.lib().grb_fog_accumulate(C.c_void_p(dev.data_ptr()), w, h, d, C.c_void_p(fog.data_ptr()), capi.stream_ptr()), "grb_fog_accumulate")
    torch.cuda.synchronize()
    got, ref = harness.to_host(fog, np.uint16), oracle.fog_accumulate(light)
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1 and (diff == 0).mean() > 0.999, (int(diff.max()), float((diff == 0).mean()))
