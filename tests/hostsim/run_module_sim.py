"""TEST INFRASTRUCTURE ONLY: run a module's __main__ (or `module:function`) with the binding
pointed at the device stand-in and torch.cuda patched to report a device, to check the host
control flow of entry points (acg_b200.driver, __graft_entry__.smoke) without a GPU."""
import importlib
import os
import runpy
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

if __name__ == "__main__":
    target = sys.argv[1]
    sys.argv = [target] + sys.argv[2:]
    if ":" in target:
        mod, fn = target.split(":")
        getattr(importlib.import_module(mod), fn)()
    else:
        runpy.run_module(target, run_name="__main__")
