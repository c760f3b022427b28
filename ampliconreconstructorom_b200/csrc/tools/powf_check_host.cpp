// Host-side exhaustive check of the powf12 restatement (powf12.cuh compiled as plain C++) against the
// libm powf the reference links.  usage: powf_check_host [lo_bits hi_bits] [nthreads]
// Compile WITHOUT -ffast-math; -mfma makes fma() a hardware instruction (else libm's exact fma is used).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <thread>
#include <vector>
#include <atomic>
#include "../powf12.cuh"
static const sa::PowLogEntry LOGTAB[16] = {SA_POWF_LOG2_TAB};
static const uint64_t EXPTAB[32] = {SA_EXP2F_TAB};
int main(int argc, char **argv) {
    uint64_t lo = argc > 2 ? strtoull(argv[1], 0, 0) : 0, hi = argc > 2 ? strtoull(argv[2], 0, 0) : 0x7f800001ull;
    int nt = argc > 3 ? atoi(argv[3]) : (int)std::thread::hardware_concurrency();
    std::atomic<uint64_t> bad{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t]() {
            uint64_t nbad = 0;
            for (uint64_t u = lo + t; u < hi; u += nt) {
                float x = sa::u2f((uint32_t)u);
                volatile float xv = x;
                float ref = powf(xv, 1.2f);
                float got = sa::powf12_generic(x, LOGTAB, EXPTAB);
                if (sa::f2u(ref) != sa::f2u(got) && !(ref != ref && got != got)) {
                    if (nbad < 5) fprintf(stderr, "MISMATCH x=%a (0x%08x) ref=%a got=%a\n", x, (unsigned)u, ref, got);
                    nbad++;
                }
            }
            bad += nbad;
        });
    for (auto &x : th) x.join();
    printf("checked [0x%llx,0x%llx) mismatches=%llu\n", (unsigned long long)lo, (unsigned long long)hi,
           (unsigned long long)bad.load());
    return bad.load() ? 1 : 0;
}
