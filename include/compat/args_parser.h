/* forwards to the B200 replacement of the GASAL2 API (see gasal_b200_compat.h) */
#include "gasal_b200_compat.h"
