/* the reference includes "./GPUSeed/seed_gen.h" (src/fastmap.c:15, src/bwamem.h:7): with
 * -I include/compat ahead of its own src/ directory that path resolves here */
#include "../seed_gen.h"
