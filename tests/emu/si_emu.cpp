// TEST INFRASTRUCTURE ONLY -- host emulator of the stage machine.
//
// Executes the programs produced by sympleints_b200/csrc/si_program.cpp with the
// same executor/setup code the CUDA kernels use (si_machine.h, compiled for the
// host, one "lane").  It exists so that the program builder and executor logic
// can be checked against the oracle in a container without a GPU
// (`pytest -m "not gpu"`).  It is never loaded by the product package.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "../../sympleints_b200/csrc/si_machine.h"
#include "../../sympleints_b200/csrc/si_program.h"

namespace {
struct Cached {
    SiProgramHost ph;
    SiProg prog;
};
std::map<std::tuple<int, int, int, int, int, int>, Cached*> g_cache;

Cached* get_prog(int key, int La, int Lb, int Lc, int sph, int normalize) {
    auto k = std::make_tuple(key, La, Lb, Lc, sph, normalize);
    auto it = g_cache.find(k);
    if (it != g_cache.end()) return it->second;
    Cached* c = new Cached;
    c->ph = si_build_program(key, La, Lb, Lc, sph, normalize);
    c->prog = c->ph.desc;
    c->prog.hdr = c->ph.hdr.data();
    c->prog.terms = c->ph.terms.data();
    c->prog.stages = c->ph.stages.data();
    c->prog.consts = c->ph.consts.data();
    g_cache[k] = c;
    return c;
}
}  // namespace

extern "C" {

// Program statistics: out[0]=w_size, [1]=nstages, [2]=n_fma_prim, [3]=n_fma_contr,
// [4]=lut words, [5]=nconst, [6]=chunk, [7]=x_size
int siemu_program_stats(int key, int La, int Lb, int Lc, int sph, int normalize, long long* out) {
    try {
        Cached* c = get_prog(key, La, Lb, Lc, sph, normalize);
        out[0] = c->prog.w_size;
        out[1] = c->prog.nstages;
        out[2] = c->ph.n_fma_prim;
        out[3] = c->ph.n_fma_contr;
        out[4] = (long long)(c->ph.hdr.size() + c->ph.terms.size());
        out[5] = c->prog.nconst;
        out[6] = c->prog.chunk;
        out[7] = c->prog.x_size;
        return 0;
    } catch (std::exception& e) {
        fprintf(stderr, "siemu_program_stats: %s\n", e.what());
        return 1;
    }
}

int siemu_cart2sph(int l, double* out) {
    std::vector<double> C = si_cart2sph(l);
    memcpy(out, C.data(), C.size() * sizeof(double));
    return 0;
}

// One contracted shell block in the reference layout (comp, a, b[, c]); La >= Lb.
// For coul, ncharge unit... charges: R points to ncharge*3 coordinates, q to ncharge weights
// (result = sum_C q_C * block(R_C)); for the multipole keys R is the origin (ncharge ignored).
int siemu_eval(int key, int La, int Lb, int Lc, int sph, int normalize, int Ka, const double* ax,
               const double* da, const double* A, int Kb, const double* bx, const double* db,
               const double* B, int Kc, const double* cx, const double* dc, const double* C,
               const double* R, int ncharge, const double* q, const double* boys_table, int nrows,
               int ncols, const double* xlarge_pref, double* result) {
    try {
        Cached* c = get_prog(key, La, Lb, Lc, sph, normalize);
        const SiProg& P = c->prog;
        SiBoys Bt;
        Bt.table = boys_table;
        Bt.tableT = nullptr;
        Bt.xlarge_pref = xlarge_pref;
        Bt.nrows = nrows;
        Bt.ncols = ncols;
        std::vector<double> W(P.w_size, 0.0), S(SI_NDYN + P.nconst, 0.0);
        for (int i = 0; i < P.nconst; ++i) S[SI_NDYN + i] = P.consts[i];
        auto nosync = []() {};
        for (int i = 0; i < P.x_size; ++i) W[P.x_off + i] = 0.0;
        for (int i = 0; i < Ka; ++i)
            for (int j = 0; j < Kb; ++j) {
                if (key == SI_KEY_3C2E) {
                    for (int k = 0; k < Kc; ++k) {
                        si_setup_3c2e(P, Bt, ax[i], da[i], A, bx[j], db[j], B, cx[k], dc[k], C, W.data(),
                                      S.data(), 0, 1);
                        si_run_stages(P, P.stages, 0, P.n_prim_stages, W.data(), S.data(), 0, 1, 0, 1, nosync);
                    }
                } else if (key == SI_KEY_COUL) {
                    for (int k = 0; k < ncharge; ++k) {
                        si_setup_coul(P, Bt, ax[i], da[i], A, bx[j], db[j], B, R + 3 * k, q[k], W.data(),
                                      S.data(), 0, 1);
                        si_run_stages(P, P.stages, 0, P.n_prim_stages, W.data(), S.data(), 0, 1, 0, 1, nosync);
                    }
                } else if (key == SI_KEY_2C2E) {
                    si_setup_2c2e(P, Bt, ax[i], da[i], A, bx[j], db[j], B, W.data(), S.data(), 0, 1);
                    si_run_stages(P, P.stages, 0, P.n_prim_stages, W.data(), S.data(), 0, 1, 0, 1, nosync);
                } else {
                    si_setup_1e(P, ax[i], da[i], A, bx[j], db[j], B, R, W.data(), S.data(), 0, 1);
                    si_run_stages(P, P.stages, 0, P.n_prim_stages, W.data(), S.data(), 0, 1, 0, 1, nosync);
                }
            }
        // geometry-only scalars needed by the contracted phase (AB for the HRR)
        {
            double v[12] = {0}, kb[5] = {0};
            for (int d = 0; d < 3; ++d) v[6 + d] = A[d] - B[d];
            si_fill_scalars(S.data(), v, kb, 0.0, 0, 1);
        }
        si_run_stages(P, P.stages, P.n_prim_stages, P.chunk_begin, W.data(), S.data(), 0, 1, 0, 1, nosync);
        for (int m0 = 0; m0 < P.nb_total; m0 += P.chunk) {
            int nb = P.nb_total - m0 < P.chunk ? P.nb_total - m0 : P.chunk;
            si_run_stages(P, P.stages, P.chunk_begin, P.nstages, W.data(), S.data(), m0, nb, 0, 1, nosync);
        }
        int n = P.out_n0 * P.out_n1 * P.nb_total;
        for (int i = 0; i < n; ++i) result[i] = W[P.out_off + i];
        return 0;
    } catch (std::exception& e) {
        fprintf(stderr, "siemu_eval: %s\n", e.what());
        return 1;
    }
}
}
