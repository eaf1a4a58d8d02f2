#define NMPC_EQ true
#define NMPC_SB true
#define NMPC_LIN true
#define NMPC_LAUNCHER launch_lin_11
#include "nmpc_variant.inc"
