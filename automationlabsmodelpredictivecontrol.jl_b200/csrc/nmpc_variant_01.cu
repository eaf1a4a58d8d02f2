#define NMPC_EQ false
#define NMPC_SB true
#define NMPC_LAUNCHER launch_sqp_01
#include "nmpc_variant.inc"
