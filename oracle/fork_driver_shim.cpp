/* oracle/fork_driver_shim.cpp -- TEST INFRASTRUCTURE ONLY.
 * The fork's SHD pre-filter (src/vector_filter.h:26, used by mem_shd_flt_chained_seeds, src/bwamem.c:889,950)
 * needs Boost.Preprocessor (src/mask.h:5), which this image does not have.  opt->shd_filter is 0 by default
 * (src/bwamem.c:143), so the filter is never reached; this stub only satisfies the linker and aborts if called. */
#include <stdio.h>
#include <stdlib.h>
extern "C" int bit_vec_filter_sse1(char *read, char *ref, int length, int max_error)
{
    (void)read; (void)ref; (void)length; (void)max_error;
    fprintf(stderr, "[fork_driver_shim] the SHD filter is not built in this test binary\n");
    abort();
}
