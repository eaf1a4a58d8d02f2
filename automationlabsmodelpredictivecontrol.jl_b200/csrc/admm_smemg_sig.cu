// sigma > 0 instantiations of the shared-memory resident general-row kernel (see admm_smemg_launch.inc)
#define MPCB_SMEMG_SIG 1
#include "admm_smemg_launch.inc"
