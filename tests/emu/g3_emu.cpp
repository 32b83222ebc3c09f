// TEST INFRASTRUCTURE ONLY -- warp emulation of the GENERATED 3c2e kernels on host threads.
// 32 std::threads play the lanes of one warp; __syncwarp is a std::barrier.  The generated source
// (tests/emu/g3_emu_classes.inc, written by codegen/gen3c2e.py:write_emu_source) is the same text
// nvcc compiles for the GPU.
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "g3_emu_classes.inc"

thread_local SiG3EmuWarp* si_g3_emu_warp = nullptr;
thread_local int si_g3_emu_lane = 0;
thread_local double* si_g3_emu_smem = nullptr;

extern "C" int g3emu_eval(int La, int Lb, int Lc, int Ka, const double* ax, const double* da, const double* A_,
                          int Kb, const double* bx, const double* db, const double* B, int naux, const int* Kc,
                          const double* cx, const double* dc, const double* C /* naux*3 */,
                          const double* boys_table, int nrows, int ncols, const double* xlarge_pref, int use_table,
                          double* result /* naux blocks, each (na, nb, nc) */) {
    const int id = La * 100 + Lb * 10 + Lc;
    const G3EmuClass* cls = nullptr;
    for (int i = 0; i < g3_emu_nclasses; ++i)
        if (g3_emu_classes[i].id == id) cls = &g3_emu_classes[i];
    if (!cls) return 2;
    // orbital "basis": shells 0 (a) and 1 (b); aux basis: naux shells of angular momentum Lc
    int oL[2] = {La, Lb}, onp[2] = {Ka, Kb}, opoff[2] = {0, Ka}, ofoff[2] = {0, 2 * La + 1};
    std::vector<double> ocen(6), oex, oco;
    for (int d = 0; d < 3; ++d) {
        ocen[d] = A_[d];
        ocen[3 + d] = B[d];
    }
    oex.insert(oex.end(), ax, ax + Ka);
    oex.insert(oex.end(), bx, bx + Kb);
    oco.insert(oco.end(), da, da + Ka);
    oco.insert(oco.end(), db, db + Kb);
    std::vector<int> aL(naux, Lc), anp(Kc, Kc + naux), apoff(naux), afoff(naux), alist(naux);
    int np = 0;
    for (int i = 0; i < naux; ++i) {
        apoff[i] = np;
        np += Kc[i];
        afoff[i] = i * (2 * Lc + 1);
        alist[i] = i;
    }
    std::vector<double> oexp(np);
    for (int i = 0; i < np; ++i) oexp[i] = 1.0 / cx[i];
    G3Args args;
    memset(&args, 0, sizeof(args));
    args.aux_oexp = oexp.data();
    args.orb = BasisDev{oL, onp, opoff, ofoff, ocen.data(), oex.data(), oco.data()};
    args.aux = BasisDev{aL.data(), anp.data(), apoff.data(), afoff.data(), C, cx, dc};
    // use_table == 2: a second pair (a shifted by (0.5, -0.25, 0.125), b) whose tuples are interleaved with the
    // first pair's in the tuple table (tuples of different pairs inside one warp)
    const int npairs = use_table == 2 ? 2 : 1;
    const double A2[3] = {A_[0] + 0.5, A_[1] - 0.25, A_[2] + 0.125};
    SiPairItem items[2];
    memset(items, 0, sizeof(items));
    SiPairItem& item = items[0];
    item.sa = 0;
    item.sb = 1;
    std::vector<double> pp((size_t)npairs * Ka * Kb * SI_PP_STRIDE, 0.0);
    for (int ipair = 0; ipair < npairs; ++ipair)
    for (int i = 0; i < Ka; ++i)
        for (int j = 0; j < Kb; ++j) {
            const double* A = ipair ? A2 : A_;
            double* q = &pp[((size_t)ipair * Ka * Kb + (size_t)(i * Kb + j)) * SI_PP_STRIDE];
            double p = ax[i] + bx[j], oop = 1.0 / p, mu = ax[i] * bx[j] * oop, r2 = 0.0;
            q[0] = p;
            q[1] = oop;
            for (int d = 0; d < 3; ++d) {
                q[2 + d] = (ax[i] * A[d] + bx[j] * B[d]) * oop;
                q[5 + d] = q[2 + d] - A[d];
                r2 += (A[d] - B[d]) * (A[d] - B[d]);
            }
            q[8] = exp(-mu * r2) * da[i] * db[j];
        }
    item.ppoff = 0;
    item.Ka = Ka;
    item.Kb = Kb;
    item.fa0 = 0;
    item.fb0 = 2 * La + 1;
    for (int d = 0; d < 3; ++d) item.AB[d] = A_[d] - B[d];
    items[1] = item;
    items[1].sa = 2;
    items[1].ppoff = Ka * Kb;
    items[1].fa0 = 2 * La + 1 + 2 * Lb + 1;
    for (int d = 0; d < 3; ++d) items[1].AB[d] = A2[d] - B[d];
    std::vector<SiAuxItem> aitems(naux);
    for (int i = 0; i < naux; ++i) {
        aitems[i].shell = i;
        aitems[i].Kc = Kc[i];
        aitems[i].pc0 = apoff[i];
        aitems[i].foff = afoff[i];
        for (int d = 0; d < 3; ++d) aitems[i].C[d] = C[3 * i + d];
    }
    args.auxitems = aitems.data();
    args.ppair = pp.data();
    args.pairs = items;
    args.npairs = npairs;
    args.auxlist = alist.data();
    args.naux_c = naux;
    args.col_begin = 0;
    args.groups_per_pair = (naux + cls->tpw - 1) / cls->tpw;
    args.ngroups = args.groups_per_pair;
    // explicit tuple table (use_table != 0): the aux shells in reverse order, one padding slot per group
    std::vector<int2> table;
    if (use_table && (cls->tpw > 1 || use_table == 2)) {
        int k = naux - 1, pend = 0;
        while (k >= 0) {
            for (int t = 0; t < cls->tpw; ++t) {
                if ((cls->tpw > 1 && t == cls->tpw - 1) || k < 0) {
                    table.push_back(int2{-1, -1});
                } else if (npairs == 2) {
                    table.push_back(int2{pend, k});
                    if (pend == 1) --k;
                    pend ^= 1;
                } else {
                    table.push_back(int2{0, k--});
                }
            }
        }
        args.tuples = table.data();
        args.ngroups = (long long)(table.size() / cls->tpw);
    }
    args.chunk = 2;
    args.screen_eps2 = 0.0;
    const int na = 2 * La + 1, nb = 2 * Lb + 1, nc = 2 * Lc + 1;
    // full layout of a 2-shell orbital basis: (nbf, nbf, naux*nc); we extract the (a, b) block afterwards
    const int nbf = na + nb + (npairs == 2 ? na : 0);
    std::vector<double> full((size_t)nbf * nbf * naux * nc, 0.0);
    args.od.out = full.data();
    args.od.ld = (long long)naux * nc;
    args.od.nbf = nbf;
    args.od.mode = STORE_3C_FULL;
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
            (args.tuples ? cls->tbody : cls->body)(args, lane, smem.data());
        });
    for (auto& t : th) t.join();
    for (int k = 0; k < naux; ++k)
        for (int i = 0; i < na; ++i)
            for (int j = 0; j < nb; ++j)
                for (int m = 0; m < nc; ++m)
                    result[(((size_t)k * na + i) * nb + j) * nc + m] =
                        full[((size_t)i * nbf + (na + j)) * args.od.ld + (size_t)k * nc + m];
    if (npairs == 2)
        for (int k = 0; k < naux; ++k)
            for (int i = 0; i < na; ++i)
                for (int j = 0; j < nb; ++j)
                    for (int m = 0; m < nc; ++m)
                        result[(((size_t)(naux + k) * na + i) * nb + j) * nc + m] =
                            full[((size_t)(na + nb + i) * nbf + (na + j)) * args.od.ld + (size_t)k * nc + m];
    return 0;
}
