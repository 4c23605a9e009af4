"""vipcup_b200 -- B200 (sm_100a) implementation of the vip-cup-2022 inference hot path.

Host side of the C-ABI library ``libvipcup.so`` (include/vipcup.h).  Mirrors the reference's Python seams:
``dataset.build_dataset`` (dataset/dataset.py:64-102), ``Config`` (utils/config.py:4-6), ``get_device``
(utils/device.py:3-13) and ``main.predict_soln`` (main.py:58-149).  PyTorch is used for device memory, streams and
torch.distributed only.  There is no CPU fallback: importing works anywhere, computing needs the CUDA library and a GPU.
"""
import os as _os

__version__ = "0.1.0"
PACKAGE_DIR = globals().get("_real") or _os.path.dirname(_os.path.abspath(__file__))
