// Host build of the per-read region-finishing source (bwa-mem_gpu_b200/csrc/region_core.cuh).
// TEST INFRASTRUCTURE: lets the exact device source run against the oracle and the fork's golden vectors on the CPU box
// (tests/test_region_host.py); never shipped and never used as a compute path.
//   g++ -O2 -shared -fPIC -I bwa-mem_gpu_b200/csrc tests/host_emul/region_host.cpp -o tests/host_emul/libregion_host.so
#include <vector>
#include "region_core.cuh"

using namespace b200region;

extern "C" int region_host_read(const Opt *o, int64_t l_pac, const int32_t *ctg_alt, const uint8_t *fwd, int l_query, const uint8_t *query,
                                int n, Reg *a, int64_t id, int *n_pri)
{
    std::vector<EH> eh((size_t)l_query + 2);
    std::vector<int32_t> z((size_t)(n > 0 ? n : 1));
    return finish_read(*o, l_pac, ctg_alt, ByteRef{fwd}, ByteQuery{query}, n, a, id, n_pri, eh.data(), z.data());
}

// the sorts alone; which = 0 end, 1 score, 2 hash, 3 hash2
extern "C" void region_host_sort(int comb, int which, int n, Reg *a)
{
    if (n <= 0) return;
    if (comb) {
        if (which == 0) combsort(LtEnd(), n, a); else if (which == 1) combsort(LtScore(), n, a);
        else if (which == 2) combsort(LtHash(), n, a); else combsort(LtHash2(), n, a);
    } else {
        if (which == 0) introsort(LtEnd(), n, a); else if (which == 1) introsort(LtScore(), n, a);
        else if (which == 2) introsort(LtHash(), n, a); else introsort(LtHash2(), n, a);
    }
}
