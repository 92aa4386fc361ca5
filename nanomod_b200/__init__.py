"""nanomod_b200 -- B200-native per-position two-sample testing stage of NanoMod ``detect``.

Public API:
  Pileup, synthetic_pileup           CSR form of the reference's per-position signal lists
  DetectOptions                      the reference's ``detect`` options (same names, defaults)
  Detector                           one GPU's engine: detect() (host buffers) / detect_device()
  SignTestTable                      the per-position table: ranked(), called_sites(), save_test()
  myDetect                           drop-in mirror of the reference seam
                                     (mfilter_coverage / mtest2 / getKStest / save_test)
  packer                             per-read event tables -> Pileup (ReadRecord, ReadFilter,
                                     pack_reads; the reference's ReadAllFast5 loop)
The arithmetic runs only in the CUDA library (nanomod_b200/_C/libnanomod_b200.so, built by
``python -m nanomod_b200.build``); importing this package does not need a GPU, calling it does.
"""
from .pileup import Pileup, synthetic_pileup, planted_shift, SYN_SEED
from .detect import (DetectOptions, Detector, DevicePileup, SignTestTable, OptionError,
                     alloc_device_table)
from ._lib import NmError, LIB_PATH
from . import packer

__all__ = ["Pileup", "synthetic_pileup", "planted_shift", "SYN_SEED", "DetectOptions", "Detector",
           "DevicePileup", "SignTestTable", "OptionError", "alloc_device_table", "NmError",
           "LIB_PATH", "packer"]
__version__ = "0.1.0"
