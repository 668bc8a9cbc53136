"""statespacer_b200 -- B200-native (sm_100a, FP64) Kalman hot path of the R package statespacer.

`api` mirrors the reference's five native routines (R/RcppExports.R:11-99) and adds the batched
loglikelihood; `sysmat` builds system matrices in the GetSysMat/CombineSysMat layout; `native` is the
ctypes binding of the C-ABI library (include/ssgpu.h).  Nothing here falls back to the CPU.
"""
__version__ = "0.1.0"
