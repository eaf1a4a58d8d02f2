#define NMPC_EQ true
#define NMPC_SB false
#define NMPC_LAUNCHER launch_sqp_10
#include "nmpc_variant.inc"
