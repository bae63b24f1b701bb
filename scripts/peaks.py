"""Measured FMA ceilings of the device (noc_measure_fma_peak): chain peak (fp32/fp64) and the 8x8 register-tile
inner loop at 8 and 16 resident warps per SM."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuraloc_b200 as nb
lib = nb._cabi.lib()
for code, name in ((0, "fp32 FMA chain, full occupancy"), (1, "fp64 FMA chain, full occupancy"),
                   (2, "fp32 8x8 register-tile inner loop, 8 warps/SM"), (3, "fp32 8x8 register-tile inner loop, 16 warps/SM"),
                   (4, "fp32 packed FFMA2 chain, full occupancy"), (5, "fp32 8x8 tile inner loop with FFMA2, 8 warps/SM"),
                   (6, "fp32 8x8 tile inner loop with FFMA2, 16 warps/SM")):
    v = C.c_double(0)
    nb._cabi.check(lib.noc_measure_fma_peak(code, v))
    print("%-55s %.2f TFLOP/s" % (name, v.value))
