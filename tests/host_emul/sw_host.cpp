// Host build of the ksw_align2 replay (bwa-mem_gpu_b200/csrc/sw_core.cuh).  TEST INFRASTRUCTURE: lets the exact kernel source be
// checked against the oracle and the reference on the CPU box (tests/test_sw_host.py); never shipped, never a compute path.
//   g++ -O2 -shared -fPIC -I include -I bwa-mem_gpu_b200/csrc tests/host_emul/sw_host.cpp -o tests/host_emul/libsw_host.so
#include <stdint.h>
#include <string.h>
#include <vector>
#include "sw_core.cuh"

extern "C" void sw_host_run(const bwa_b200_ext_params_t *p, uint64_t n, const uint8_t *qseq, const uint32_t *qoff, const uint32_t *qlen,
                            const uint8_t *tseq, const uint32_t *toff, const uint32_t *tlen, const uint32_t *xtra, bwa_b200_sw_result_t *out)
{
    SwParams S;
    memset(&S, 0, sizeof(S));
    memcpy(S.mat, p->mat, 25);
    S.m = 5; S.o_del = p->o_del; S.e_del = p->e_del; S.o_ins = p->o_ins; S.e_ins = p->e_ins;
    std::vector<int16_t> ws, rm;
    for (uint64_t a = 0; a < n; ++a) {
        const size_t n_cap = ((size_t)qlen[a] + 15) / 16 * 16 + 16;
        ws.assign(4 * n_cap, 0); rm.assign((size_t)tlen[a] + 1, 0);
        sw_align2((int)qlen[a], qseq + qoff[a], (int)tlen[a], tseq + toff[a], S, (int)xtra[a], ws.data(), 1, n_cap, rm.data(), out[a]);
    }
}
