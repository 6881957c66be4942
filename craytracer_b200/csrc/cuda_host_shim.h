// Host translation units include cray_math.cuh without nvcc: nothing to shim beyond the CRAY_HD macro it defines itself.
#pragma once
