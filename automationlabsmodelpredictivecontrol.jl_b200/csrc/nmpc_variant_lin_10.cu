#define NMPC_EQ true
#define NMPC_SB false
#define NMPC_LIN true
#define NMPC_LAUNCHER launch_lin_10
#include "nmpc_variant.inc"
