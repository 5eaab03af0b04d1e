"""TEST INFRASTRUCTURE ONLY: run bench.py's main() unchanged on the device stand-in
(tests/hostsim) to check its control flow and the shape of the JSON line it prints --
set-up of the matrix for 1 and N ranks, timed pass, profiled pass, roofline / e2e /
cpu_baseline fields -- on a machine without a GPU.  Nothing it prints is a measurement.
torch.cuda is patched just enough for bench.py to proceed (availability, synchronize);
the nvidia-smi sampler is replaced by a stub.  bench.py itself has no such switch."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import torch                                          # noqa: E402

import acg_b200.api as api                            # noqa: E402
api._LIBPATH = os.path.join(HERE, "libacgb200_hostsim.so")
torch.cuda.is_available = lambda: True
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.device_count = lambda: 8

import bench                                          # noqa: E402


class _NoSampler:
    def __init__(self, index=0):
        pass

    def start(self):
        pass

    def stop(self, t0, t1):
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "host simulation: no device to sample"}


bench.ClockSampler = _NoSampler

if __name__ == "__main__":
    sys.argv = ["bench.py"] + sys.argv[1:]
    sys.exit(bench.main())
