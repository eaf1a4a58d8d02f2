#define NMPC_EQ false
#define NMPC_SB false
#define NMPC_LAUNCHER launch_sqp_00
#include "nmpc_variant.inc"
