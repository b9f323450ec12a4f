/* oracle/ref_shim -- TEST INFRASTRUCTURE: mtk/startIdx.hpp (subblock / setDiagonal live in src/SubManifold.hpp here) */
#ifndef REF_SHIM_MTK_STARTIDX
#define REF_SHIM_MTK_STARTIDX
#include "src/SubManifold.hpp"
#endif
