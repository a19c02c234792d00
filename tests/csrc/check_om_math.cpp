// CPU check of oscen_b200/csrc/om_math.h against the platform libm (glibc).
// usage: check_om_math <stride> [threads]   (stride 1 = every f32 bit pattern)
// prints one line per function: name n_checked n_mismatch max_ulp
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include "../../oscen_b200/csrc/om_math.h"

static inline float asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int64_t ord(float f) { int32_t i; memcpy(&i, &f, 4); return i < 0 ? (int64_t)0x80000000LL - (int64_t)(uint32_t)i - 0x80000000LL + 0 : i; }
static inline int64_t ulpdiff(float a, float b) {
    if (a == b) return 0;
    if (std::isnan(a) && std::isnan(b)) return 0;
    int32_t ia, ib; memcpy(&ia, &a, 4); memcpy(&ib, &b, 4);
    int64_t oa = ia < 0 ? -(int64_t)(ia & 0x7fffffff) : ia;
    int64_t ob = ib < 0 ? -(int64_t)(ib & 0x7fffffff) : ib;
    return oa > ob ? oa - ob : ob - oa;
}
struct Res { uint64_t n = 0, bad = 0; int64_t maxulp = 0; uint32_t worst = 0; };
template <class F, class G>
static Res sweep(F mine, G ref, uint32_t lo_bits, uint32_t hi_bits, uint32_t stride, int nt) {
    std::vector<Res> rs(nt);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([&, t] {
        Res r;
        for (uint64_t u = (uint64_t)lo_bits + (uint64_t)t * stride; u <= hi_bits; u += (uint64_t)stride * nt) {
            for (int sgn = 0; sgn < 2; ++sgn) {
                float x = asfloat((uint32_t)u | (sgn ? 0x80000000u : 0u));
                float a = mine(x), b = ref(x);
                int64_t d = ulpdiff(a, b);
                r.n++;
                if (d) { r.bad++; if (d > r.maxulp) { r.maxulp = d; r.worst = (uint32_t)u | (sgn ? 0x80000000u : 0u); } }
            }
        }
        rs[t] = r;
    });
    for (auto &x : th) x.join();
    Res tot;
    for (auto &r : rs) { tot.n += r.n; tot.bad += r.bad; if (r.maxulp > tot.maxulp) { tot.maxulp = r.maxulp; tot.worst = r.worst; } }
    return tot;
}
static void report(const char *name, Res r) {
    printf("%s n=%llu mismatch=%llu max_ulp=%lld worst=0x%08x\n", name, (unsigned long long)r.n,
           (unsigned long long)r.bad, (long long)r.maxulp, r.worst);
    fflush(stdout);
}
int main(int argc, char **argv) {
    uint32_t stride = argc > 1 ? (uint32_t)strtoul(argv[1], 0, 10) : 1009;
    int nt = argc > 2 ? atoi(argv[2]) : 8;
    const uint32_t F104 = 0x42d00000u /*104*/, F120 = 0x42f00000u, FINF = 0x7f800000u;
    report("expf[|x|<=104]", sweep([](float x) { return om_expf(x); }, [](float x) { return expf(x); }, 0, F104, stride, nt));
    report("sinf[|x|<120]", sweep([](float x) { return om_sinf(x); }, [](float x) { return sinf(x); }, 0, F120 - 1, stride, nt));
    report("cosf[|x|<120]", sweep([](float x) { return om_cosf(x); }, [](float x) { return cosf(x); }, 0, F120 - 1, stride, nt));
    report("sinf_u[all]", sweep([](float x) { return om_sinf_u(x); }, [](float x) { return sinf(x); }, 0, FINF - 1, stride, nt));
    report("cosf_u[all]", sweep([](float x) { return om_cosf_u(x); }, [](float x) { return cosf(x); }, 0, FINF - 1, stride, nt));
    report("sinf_r[|x|<120] (both polynomials, no range test)", sweep([](float x) { return om_sinf_r(x); }, [](float x) { return sinf(x); }, 0, F120 - 1, stride, nt));
    // k_tanf_mid: only positive x in [2^-13, 0.6744) — compare against the general routine on that range
    report("tanf_mid[2^-13<=x<0.6744]", sweep([](float x) { return x > 0.0f ? om_k_tanf_mid(x) : om_tanf(x); }, [](float x) { return om_tanf(x); }, 0x39000000u, 0x3f2ca140u - 1u, stride, nt));
    report("sinf[120<=|x|<inf]", sweep([](float x) { return om_sinf(x); }, [](float x) { return sinf(x); }, F120, FINF - 1, stride, nt));
    report("cosf[120<=|x|<inf]", sweep([](float x) { return om_cosf(x); }, [](float x) { return cosf(x); }, F120, FINF - 1, stride, nt));
    const uint32_t F2 = 0x40000000u, F12 = 0x41400000u;
    report("tanf[|x|<=2]", sweep([](float x) { return om_tanf(x); }, [](float x) { return tanf(x); }, 0, F2, stride, nt));
    report("tanhf_u[|x|<=32]", sweep([](float x) { return om_tanhf_u(x); }, [](float x) { return tanhf(x); }, 0, 0x42000000u, stride, nt));
    report("tanhf[|x|<=12]", sweep([](float x) { return om_tanhf(x); }, [](float x) { return tanhf(x); }, 0, F12, stride, nt));
    // ADSR domain: 1 - exp(-K/n) for integer n (adsr.rs:132-133)
    {
        uint64_t bad = 0, n = 0;
        for (uint32_t k = 1; k < (1u << 24); ++k) {
            float x = -4.6051702f / (float)k;
            if (om_expf(x) != expf(x)) ++bad;
            ++n;
        }
        printf("expf[adsr -K/n, n<2^24] n=%llu mismatch=%llu\n", (unsigned long long)n, (unsigned long long)bad);
    }
    return 0;
}
