#define NMPC_EQ false
#define NMPC_SB true
#define NMPC_LIN true
#define NMPC_LAUNCHER launch_lin_01
#include "nmpc_variant.inc"
