#define NMPC_EQ false
#define NMPC_SB false
#define NMPC_LIN true
#define NMPC_LAUNCHER launch_lin_00
#include "nmpc_variant.inc"
