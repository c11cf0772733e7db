// register-tiled MultiStateAligner11ts kernel, 6 columns per lane (see msa_tiled.cuh)
#include "msa_kernels.cuh"
BBM_DEFINE_TILED_LAUNCH(6)
