// TEST INFRASTRUCTURE ONLY -- warp emulation of the GENERATED nuclear-attraction kernels on host threads
// (see g3_emu.cpp).  The generated source (tests/emu/gc_emu_classes.inc, written by
// codegen/gencoul.py:write_emu_source) is the same text nvcc compiles for the GPU.
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "gc_emu_classes.inc"

thread_local SiG3EmuWarp* si_g3_emu_warp = nullptr;
thread_local int si_g3_emu_lane = 0;
thread_local double* si_g3_emu_smem = nullptr;

// npairs shell pairs of one class (La >= Lb); pair i has Ka[i] x Kb[i] primitives (concatenated arrays).
// result: npairs blocks (nsa, nsb) = sum_C q_C (a|1/r_C|b).
extern "C" int gcemu_eval(int La, int Lb, int npairs, const int* Ka, const double* ax, const double* da,
                          const double* A /* npairs*3 */, const int* Kb, const double* bx, const double* db,
                          const double* B, int ncharge, const double* cxyz, const double* cq,
                          const double* boys_table, int nrows, int ncols, const double* xlarge_pref,
                          int nsplit, double* result) {
    const int id = La * 10 + Lb;
    const GCEmuClass* cls = nullptr;
    for (int i = 0; i < gc_emu_nclasses; ++i)
        if (gc_emu_classes[i].id == id) cls = &gc_emu_classes[i];
    if (!cls) return 2;
    const int na = 2 * La + 1, nb = 2 * Lb + 1;
    const int nbf = npairs * (na + nb);
    std::vector<SiPairItem> items(npairs);
    std::vector<double> pp;
    int oa = 0, ob = 0;
    for (int ip = 0; ip < npairs; ++ip) {
        SiPairItem& it = items[ip];
        memset(&it, 0, sizeof(it));
        it.sa = 2 * ip;
        it.sb = 2 * ip + 1;
        it.Ka = Ka[ip];
        it.Kb = Kb[ip];
        it.fa0 = ip * (na + nb);
        it.fb0 = it.fa0 + na;
        it.ppoff = (int)(pp.size() / SI_PP_STRIDE);
        const double* Ai = A + 3 * ip;
        const double* Bi = B + 3 * ip;
        for (int d = 0; d < 3; ++d) it.AB[d] = Ai[d] - Bi[d];
        for (int i = 0; i < Ka[ip]; ++i)
            for (int j = 0; j < Kb[ip]; ++j) {
                double q[SI_PP_STRIDE] = {0};
                const double a = ax[oa + i], b = bx[ob + j];
                const double p = a + b, oop = 1.0 / p, mu = a * b * oop;
                double r2 = 0.0;
                q[0] = p;
                q[1] = oop;
                for (int d = 0; d < 3; ++d) {
                    q[2 + d] = (a * Ai[d] + b * Bi[d]) * oop;
                    q[5 + d] = q[2 + d] - Ai[d];
                    r2 += (Ai[d] - Bi[d]) * (Ai[d] - Bi[d]);
                }
                q[8] = exp(-mu * r2) * da[oa + i] * db[ob + j];
                pp.insert(pp.end(), q, q + SI_PP_STRIDE);
            }
        oa += Ka[ip];
        ob += Kb[ip];
    }
    GCArgs args;
    memset(&args, 0, sizeof(args));
    args.pairs = items.data();
    args.ppair = pp.data();
    args.npairs = npairs;
    std::vector<double> c4((size_t)4 * ncharge);
    for (int i = 0; i < ncharge; ++i) {
        for (int d = 0; d < 3; ++d) c4[4 * i + d] = cxyz[3 * i + d];
        c4[4 * i + 3] = cq[i];
    }
    args.c4 = c4.data();
    args.ncharge = ncharge;
    args.nsplit = nsplit;
    args.ngroups = (long long)((npairs + cls->tpw - 1) / cls->tpw) * nsplit;
    std::vector<double> full((size_t)nbf * nbf * nsplit, 0.0);
    args.part = full.data() + (size_t)nbf * nbf;
    args.od.out = full.data();
    args.od.ld = nbf;
    args.od.nbf = nbf;
    args.od.mode = STORE_MATRIX;
    int counter = 0;
    args.counter = &counter;
    std::vector<double> boys_T((size_t)nrows * ncols);
    for (int r = 0; r < nrows; ++r)
        for (int i = 0; i < ncols; ++i) boys_T[(size_t)i * nrows + r] = boys_table[(size_t)r * ncols + i];
    args.boys.tableT = boys_T.data();
    args.boys.table = boys_table;
    args.boys.xlarge_pref = xlarge_pref;
    args.boys.nrows = nrows;
    args.boys.ncols = ncols;
    std::vector<double> smem((size_t)cls->wsz * cls->tpw, 0.0);
    std::barrier<> bar(32);
    SiG3EmuWarp warp;
    warp.bar = &bar;
    std::vector<std::thread> th;
    for (int lane = 0; lane < 32; ++lane)
        th.emplace_back([&, lane]() {
            si_g3_emu_warp = &warp;
            si_g3_emu_lane = lane;
            si_g3_emu_smem = smem.data();
            cls->body(args, lane, smem.data(), 0, 1);
        });
    for (auto& t : th) t.join();
    for (int sp = 1; sp < nsplit; ++sp)
        for (size_t i = 0; i < (size_t)nbf * nbf; ++i) full[i] += full[(size_t)sp * nbf * nbf + i];
    for (int ip = 0; ip < npairs; ++ip)
        for (int i = 0; i < na; ++i)
            for (int j = 0; j < nb; ++j) {
                const double v = full[(size_t)(items[ip].fa0 + i) * nbf + items[ip].fb0 + j];
                const double w = full[(size_t)(items[ip].fb0 + j) * nbf + items[ip].fa0 + i];
                if (v != w) return 3;   // mirror store
                result[((size_t)ip * na + i) * nb + j] = v;
            }
    return 0;
}
