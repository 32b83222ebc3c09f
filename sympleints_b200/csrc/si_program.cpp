// Host-side program builder (see si_program.h / si_machine.h).
//
// Recurrences restated from the reference (file:line in /root/reference/sympleints):
//   defs/multipole.py:17-32   1-D multipole recurrence              -> build_1e
//   defs/kinetic.py:15-44     1-D kinetic recurrence                -> build_1e
//   defs/coulomb.py:35-80     nuclear attraction VRR (bra part)     -> build_coul
//   graphs/defs/int_nucattr.py:26-31  HRR for the ket               -> build_coul
//   defs/coulomb.py:120-160   2c2e VRR on bra and ket               -> build_2c2e
//   defs/coulomb.py:218-341   3c2e: VRR(a), truncated aux VRR(c), HRR(b)  -> build_3c2e
//   main.py:122-144,175-189,250-272  lmn factors, cart->sph on every index
//   cart2sph.py:28-88,163-202 Schlegel/Frisch coefficients
#include "si_program.h"

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>

// coul scheme: 0 (default) = the reference's two-index VRR (same conditioning as the generated
// python code); 1 = VRR on the bra + HRR (graphs/defs/int_nucattr.py; fewer flops, but the HRR
// loses digits when the ket exponent is much larger than the bra exponent).
static int g_coul_hrr = -1;
int si_coul_use_hrr() {
    if (g_coul_hrr < 0) {
        const char* e = getenv("SI_COUL_HRR");
        g_coul_hrr = (e && e[0] == '1') ? 1 : 0;
    }
    return g_coul_hrr;
}
void si_coul_set_hrr(int v) { g_coul_hrr = v ? 1 : 0; }

int si_ncart(int l) { return (l + 1) * (l + 2) / 2; }
int si_nsph(int l) { return 2 * l + 1; }
int si_ncomp(int key) {
    switch (key) {
        case SI_KEY_DPM: return 3;
        case SI_KEY_QPM: return 6;
        case SI_KEY_DQPM: return 3;
        default: return 1;
    }
}

void si_cart_order(int L, std::vector<int>& lmn) {
    lmn.clear();
    for (int i = 0; i <= L; ++i) {
        int l = L - i;
        for (int n = 0; n <= i; ++n) {
            lmn.push_back(l);
            lmn.push_back(i - n);
            lmn.push_back(n);
        }
    }
}

static long double fact_ld(int n) {
    long double r = 1.0L;
    for (int i = 2; i <= n; ++i) r *= i;
    return r;
}
static long double binom_ld(int n, int k) {
    if (k < 0 || k > n) return 0.0L;
    return fact_ld(n) / (fact_ld(k) * fact_ld(n - k));
}
static double fact2_d(int n) {
    double r = 1.0;
    while (n > 1) {
        r *= n;
        n -= 2;
    }
    return r;
}

// Complex Schlegel/Frisch coefficient, Eq. (15); returns (re, im).
static void c2s_complex(int lx, int ly, int lz, int m, long double& re, long double& im) {
    re = im = 0.0L;
    const int l = lx + ly + lz;
    const int am = std::abs(m);
    const int j2 = lx + ly - am;
    if (j2 < 0 || (j2 & 1)) return;
    const int j = j2 / 2;
    const int lmm = l - am;
    const int sign = (m >= 0) ? 1 : -1;
    long double tre = 0.0L, tim = 0.0L;
    for (int i = 0; i <= lmm / 2; ++i) {
        if (j > i) continue;
        long double ifact = binom_ld(l, i) * binom_ld(i, j) * ((i & 1) ? -1.0L : 1.0L) *
                            fact_ld(2 * l - 2 * i) / fact_ld(lmm - 2 * i);
        long double kre = 0.0L, kim = 0.0L;
        for (int k = 0; k <= j; ++k) {
            if (lx - 2 * k < 0 || lx - 2 * k > am) continue;
            int e2 = sign * (am - lx + 2 * k);  // (-1)^(e2/2) = i^e2
            int r = ((e2 % 4) + 4) % 4;
            long double w = binom_ld(j, k) * binom_ld(am, lx - 2 * k);
            if (r == 0) kre += w;
            else if (r == 1) kim += w;
            else if (r == 2) kre -= w;
            else kim -= w;
        }
        tre += ifact * kre;
        tim += ifact * kim;
    }
    long double rad = fact_ld(2 * lx) * fact_ld(2 * ly) * fact_ld(2 * lz) * fact_ld(l) * fact_ld(lmm) /
                      (fact_ld(2 * l) * fact_ld(lx) * fact_ld(ly) * fact_ld(lz) * fact_ld(l + am));
    long double f = sqrtl(rad) / (powl(2.0L, l) * fact_ld(l));
    re = tre * f;
    im = tim * f;
}

std::vector<double> si_cart2sph(int l) {
    std::vector<int> lmn;
    si_cart_order(l, lmn);
    const int nc = si_ncart(l), ns = si_nsph(l);
    std::vector<double> C((size_t)ns * nc, 0.0);
    const long double s2 = sqrtl(2.0L);
    for (int c = 0; c < nc; ++c) {
        for (int mi = 0; mi < ns; ++mi) {
            int m = mi - l;
            long double re, im;
            c2s_complex(lmn[3 * c], lmn[3 * c + 1], lmn[3 * c + 2], m, re, im);
            long double val;
            if (m == 0) {
                val = re;
            } else {
                long double re2, im2;
                c2s_complex(lmn[3 * c], lmn[3 * c + 1], lmn[3 * c + 2], -m, re2, im2);
                if (m < 0) {
                    val = (re + re2) / s2;  // (c_m + c_-m)/sqrt(2)
                } else {
                    // (c_m - c_-m)/sqrt(-2) = (x + iy)/(i sqrt2) -> real part y/sqrt2
                    val = (im - im2) / s2;
                }
            }
            double v = (double)val;
            if (std::fabs(v) <= 1e-14) v = 0.0;
            C[(size_t)mi * nc + c] = v;
        }
    }
    return C;
}

std::vector<double> si_lmn_factors(int l) {
    std::vector<int> lmn;
    si_cart_order(l, lmn);
    std::vector<double> f(si_ncart(l));
    for (int c = 0; c < si_ncart(l); ++c)
        f[c] = 1.0 / std::sqrt(fact2_d(2 * lmn[3 * c] - 1) * fact2_d(2 * lmn[3 * c + 1] - 1) *
                               fact2_d(2 * lmn[3 * c + 2] - 1));
    return f;
}

// ---------------------------------------------------------------------------
// generic DAG -> stages
// ---------------------------------------------------------------------------
namespace {

enum Region { R_TEMP = 0, R_IN = 1, R_OUT = 2 };

struct GTerm {
    int node;
    int sid;
};
struct GNode {
    std::vector<GTerm> terms;
    int level = -1;
    int addr = -1;
    int region = R_TEMP;
    bool accum = false;
    bool prod3 = false;
    int last_use = -1;
    int nref = 0;
};

struct Graph {
    std::vector<GNode> nodes;
    int add(int region = R_TEMP, int addr = -1) {
        GNode n;
        n.region = region;
        n.addr = addr;
        nodes.push_back(n);
        return (int)nodes.size() - 1;
    }
};

struct Builder {
    SiProgramHost P;
    std::map<double, int> const_ids;

    int const_sid(double v) {
        auto it = const_ids.find(v);
        if (it != const_ids.end()) return it->second;
        int id = SI_NDYN + (int)P.consts.size();
        P.consts.push_back(v);
        const_ids[v] = id;
        if (id >= 65536) throw std::runtime_error("too many constants");
        return id;
    }

    static uint32_t magic_for(int nelem, int max_total) {
        uint32_t m = (uint32_t)((((uint64_t)1) << 32) / (uint64_t)nelem) + 1u;
        if (nelem == 1) m = 0xffffffffu;  // handled specially: e/1 = e (checked below)
        for (int e = 0; e < max_total; ++e) {
            uint32_t q = (uint32_t)(((uint64_t)(uint32_t)e * m) >> 32);
            if (nelem == 1) {
                // umulhi(e, 2^32-1) = e - 1 for e>0 -> not exact; use nelem==1 fast path below
                break;
            }
            if ((int)q != e / nelem) throw std::runtime_error("magic division failed");
        }
        return m;
    }

    // Emit the stages of one DAG.  temp_off: first W address of the TEMP region.
    // batched: phase-C semantics (TEMP address = temp_off + slot, batch stride = nslots).
    // Returns the number of TEMP slots used (peak).
    int emit(Graph& G, int temp_off, bool batched, int in_bs, int in_cs, int out_bs, int out_cs,
             int max_batch, long long* fma_counter) {
        auto& N = G.nodes;
        // levels
        std::function<int(int)> level = [&](int i) -> int {
            GNode& n = N[i];
            if (n.level >= 0) return n.level;
            int lv = 0;
            for (auto& t : n.terms) lv = std::max(lv, level(t.node) + 1);
            n.level = lv;
            return lv;
        };
        int maxlevel = 0;
        for (int i = 0; i < (int)N.size(); ++i) maxlevel = std::max(maxlevel, level(i));
        for (int i = 0; i < (int)N.size(); ++i)
            for (auto& t : N[i].terms) {
                N[t.node].last_use = std::max(N[t.node].last_use, N[i].level);
                N[t.node].nref++;
            }
        std::vector<std::vector<int>> by_level(maxlevel + 1);
        for (int i = 0; i < (int)N.size(); ++i) by_level[N[i].level].push_back(i);

        std::vector<int> free_slots;
        int nslots = 0;
        std::vector<int> live;  // TEMP nodes currently holding a slot
        // stage records first, strides patched after nslots is known
        struct Pending {
            size_t stage_index;
            int src_region, dst_region;
        };
        std::vector<Pending> pend;

        for (int lv = 0; lv <= maxlevel; ++lv) {
            // release
            std::vector<int> still;
            for (int i : live) {
                if (N[i].last_use < lv) free_slots.push_back(N[i].addr);
                else still.push_back(i);
            }
            live.swap(still);
            // allocate (level-0 TEMP nodes without terms are not allowed)
            for (int i : by_level[lv]) {
                GNode& n = N[i];
                if (n.region != R_TEMP) continue;
                if (n.terms.empty()) throw std::runtime_error("TEMP node without terms");
                if (!free_slots.empty()) {
                    // take the lowest free slot to keep addresses compact
                    auto it = std::min_element(free_slots.begin(), free_slots.end());
                    n.addr = *it;
                    free_slots.erase(it);
                } else {
                    n.addr = nslots++;
                }
                live.push_back(i);
            }
            if (lv == 0) continue;
            // group by (nterms, accum, prod3, src_region, dst_region)
            std::map<std::array<int, 5>, std::vector<int>> groups;
            for (int i : by_level[lv]) {
                GNode& n = N[i];
                if (n.terms.empty()) continue;
                int sreg = -1;
                for (auto& t : n.terms) {
                    int r = N[t.node].region == R_TEMP ? R_TEMP : R_IN;
                    if (batched) {
                        if (sreg >= 0 && sreg != r) throw std::runtime_error("mixed source regions in a batched stage");
                    }
                    sreg = r;
                }
                groups[{(int)n.terms.size(), n.accum ? 1 : 0, n.prod3 ? 1 : 0, batched ? sreg : 0,
                        batched ? n.region : 0}].push_back(i);
            }
            bool first = true;
            for (auto& kv : groups) {
                auto& ids = kv.second;
                // sort by destination address for conflict-free stores
                std::sort(ids.begin(), ids.end(), [&](int x, int y) {
                    if (N[x].region != N[y].region) return N[x].region < N[y].region;
                    return N[x].addr < N[y].addr;
                });
                SiStage st;
                std::memset(&st, 0, sizeof(st));
                st.hdr_off = (int)P.hdr.size();
                st.term_off = (int)P.terms.size();
                st.nelem = (int)ids.size();
                st.nterms = kv.first[0];
                st.batched = batched ? 1 : 0;
                st.flags = (kv.first[1] ? SI_ST_ACCUM : 0) | (kv.first[2] ? SI_ST_PROD3 : 0) |
                           (first ? SI_ST_SYNC : 0);
                st.magic = magic_for(st.nelem, st.nelem * std::max(1, max_batch));
                first = false;
                for (int i : ids) {
                    GNode& n = N[i];
                    int d = n.addr + (n.region == R_TEMP ? temp_off : 0);
                    P.hdr.push_back((uint32_t)d);
                }
                P.terms.resize(P.terms.size() + (size_t)st.nterms * st.nelem);
                for (int e = 0; e < st.nelem; ++e) {
                    GNode& n = N[ids[e]];
                    for (int k = 0; k < st.nterms; ++k) {
                        const GTerm& t = n.terms[k];
                        const GNode& c = N[t.node];
                        int a = c.addr + (c.region == R_TEMP ? temp_off : 0);
                        if (a < 0 || a >= 65536) throw std::runtime_error("work array too large");
                        P.terms[st.term_off + (size_t)k * st.nelem + e] = (uint32_t)a | ((uint32_t)t.sid << 16);
                    }
                }
                pend.push_back({P.stages.size(), kv.first[3], kv.first[4]});
                P.stages.push_back(st);
                if (fma_counter) *fma_counter += (long long)st.nterms * st.nelem * std::max(1, max_batch);
            }
        }
        if (batched) {
            for (auto& pd : pend) {
                SiStage& st = P.stages[pd.stage_index];
                if (pd.src_region == R_TEMP) {
                    st.sbs = nslots;
                    st.scs = 0;
                } else {
                    st.sbs = in_bs;
                    st.scs = in_cs;
                }
                if (pd.dst_region == R_TEMP) {
                    st.dbs = nslots;
                    st.dcs = 0;
                } else {
                    st.dbs = out_bs;
                    st.dcs = out_cs;
                }
            }
        }
        return nslots;
    }
};

typedef std::array<int, 3> I3;
static int first_nz(const I3& a) {
    for (int d = 0; d < 3; ++d)
        if (a[d] > 0) return d;
    return -1;
}
static int sum3(const I3& a) { return a[0] + a[1] + a[2]; }
static I3 dec(I3 a, int d) {
    a[d] -= 1;
    return a;
}
static I3 inc(I3 a, int d) {
    a[d] += 1;
    return a;
}

struct CartIndex {
    // index of Cartesian component within its shell (canonical order)
    static int idx(const I3& a) {
        int L = sum3(a);
        int i = L - a[0];
        return i * (i + 1) / 2 + a[2];
    }
};

// list of Cartesian components for shells lo..hi, concatenated
static std::vector<I3> comps_range(int lo, int hi) {
    std::vector<I3> out;
    for (int l = lo; l <= hi; ++l) {
        std::vector<int> lmn;
        si_cart_order(l, lmn);
        for (int c = 0; c < si_ncart(l); ++c) out.push_back({lmn[3 * c], lmn[3 * c + 1], lmn[3 * c + 2]});
    }
    return out;
}
static int range_index(const I3& a, int lo) {
    int L = sum3(a);
    int off = 0;
    for (int l = lo; l < L; ++l) off += si_ncart(l);
    return off + CartIndex::idx(a);
}

// (transform matrix rows x cols) used for the final index transforms:
// sph: C2S * diag(lmn) ; cart: diag(lmn) or identity
static std::vector<double> final_transform(int l, int sph, int normalize, int& nout) {
    const int nc = si_ncart(l);
    std::vector<double> f(nc, 1.0);
    if (normalize != 0) f = si_lmn_factors(l);
    std::vector<double> M;
    if (sph) {
        nout = si_nsph(l);
        M = si_cart2sph(l);
        for (int m = 0; m < nout; ++m)
            for (int c = 0; c < nc; ++c) M[(size_t)m * nc + c] *= f[c];
    } else {
        nout = nc;
        M.assign((size_t)nc * nc, 0.0);
        for (int c = 0; c < nc; ++c) M[(size_t)c * nc + c] = f[c];
    }
    return M;
}

// Phase-C graph for 2-index-on-(a,b) transforms: optional HRR from [a'0] inputs,
// then transform of b, then of a.  Inputs are R_IN nodes created by in_node().
struct ContrSpec {
    int La, Lb;
    bool hrr;          // inputs are (a',0) with La<=|a'|<=La+Lb; else inputs are (a,b) Cartesian
    int in_off;        // W address of the input array (batch 0)
    int out_off;       // W address of the output block
    int out_estride;   // address stride between consecutive (i,j) output elements
};

static void build_contr_graph(Builder& B, Graph& G, const ContrSpec& cs, int sph, int normalize,
                              int& n0, int& n1) {
    const int La = cs.La, Lb = cs.Lb;
    std::vector<I3> ca = comps_range(La, La), cb = comps_range(Lb, Lb);
    const int nca = (int)ca.size(), ncb = (int)cb.size();
    std::vector<double> Ma = final_transform(La, sph, normalize, n0);
    std::vector<double> Mb = final_transform(Lb, sph, normalize, n1);
    // cart (a,b) nodes
    std::vector<int> ab(nca * ncb, -1);
    if (cs.hrr) {
        std::map<std::array<int, 6>, int> memo;
        std::function<int(const I3&, const I3&)> hrr = [&](const I3& a, const I3& b) -> int {
            std::array<int, 6> key = {a[0], a[1], a[2], b[0], b[1], b[2]};
            auto it = memo.find(key);
            if (it != memo.end()) return it->second;
            int id;
            if (sum3(b) == 0) {
                id = G.add(R_IN, cs.in_off + range_index(a, La));
            } else {
                int d = first_nz(b);
                int c1 = hrr(inc(a, d), dec(b, d));
                int c2 = hrr(a, dec(b, d));
                id = G.add(R_TEMP);
                G.nodes[id].terms.push_back({c1, SI_SID_ONE});
                G.nodes[id].terms.push_back({c2, SI_SID_V2 + d});  // AB_d
            }
            memo[key] = id;
            return id;
        };
        for (int i = 0; i < nca; ++i)
            for (int j = 0; j < ncb; ++j) ab[i * ncb + j] = hrr(ca[i], cb[j]);
    } else {
        for (int i = 0; i < nca; ++i)
            for (int j = 0; j < ncb; ++j) ab[i * ncb + j] = G.add(R_IN, cs.in_off + i * ncb + j);
    }
    // transform b: (a_cart, jb)
    std::vector<int> tb(nca * n1, -1);
    for (int i = 0; i < nca; ++i)
        for (int jb = 0; jb < n1; ++jb) {
            int id = G.add(R_TEMP);
            for (int j = 0; j < ncb; ++j) {
                double c = Mb[(size_t)jb * ncb + j];
                if (c != 0.0) G.nodes[id].terms.push_back({ab[i * ncb + j], B.const_sid(c)});
            }
            tb[i * n1 + jb] = id;
        }
    // transform a: (ia, jb) -> OUT
    for (int ia = 0; ia < n0; ++ia)
        for (int jb = 0; jb < n1; ++jb) {
            int id = G.add(R_OUT, cs.out_off + (ia * n1 + jb) * cs.out_estride);
            for (int i = 0; i < nca; ++i) {
                double c = Ma[(size_t)ia * nca + i];
                if (c != 0.0) G.nodes[id].terms.push_back({tb[i * n1 + jb], B.const_sid(c)});
            }
        }
}

// After building target nodes of a primitive-phase graph: make them accumulate
// into X.  Targets referenced by other nodes (or base nodes) get a copy node.
static void finalize_targets(Graph& G, const std::vector<int>& targets, int x_off) {
    std::vector<int> nref(G.nodes.size(), 0);
    for (auto& n : G.nodes)
        for (auto& t : n.terms) nref[t.node]++;
    for (size_t k = 0; k < targets.size(); ++k) {
        int id = targets[k];
        GNode& n = G.nodes[id];
        if (n.terms.empty() || nref[id] > 0 || n.region != R_TEMP) {
            int cp = G.add(R_OUT, x_off + (int)k);
            G.nodes[cp].terms.push_back({id, SI_SID_ONE});
            G.nodes[cp].accum = true;
        } else {
            n.region = R_OUT;
            n.addr = x_off + (int)k;
            n.accum = true;
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------
// key-specific builders
// ---------------------------------------------------------------------------
SiProgramHost si_build_program(int key, int La, int Lb, int Lc, int sph, int normalize) {
    if (La < Lb) throw std::runtime_error("si_build_program: requires La >= Lb");
    Builder B;
    SiProg& D = B.P.desc;
    std::memset(&D, 0, sizeof(D));
    D.La = La;
    D.Lb = Lb;
    D.Lc = Lc;
    const int L = La + Lb;

    Graph GP;  // primitive phase
    int x_size = 0, nbase = 0, nlo = 0;
    std::vector<int> targets;
    // The primitive graph addresses: X at [0, x_size), base at [x_size, x_size+nbase),
    // TEMP from x_size+nbase.  Base node addresses are patched once x_size is known,
    // so base nodes carry region R_IN with addr = base index and are fixed up below.
    std::vector<int> base_nodes;  // node ids; addr = base index for now

    bool is_1e = (key == SI_KEY_OVLP || key == SI_KEY_KIN || key == SI_KEY_DPM || key == SI_KEY_QPM ||
                  key == SI_KEY_DQPM);
    const int ncomp = si_ncomp(key);
    std::vector<I3> ca = comps_range(La, La), cb = comps_range(Lb, Lb);

    if (is_1e) {
        nbase = (key == SI_KEY_KIN) ? 6 : 3;
        // 1-D tables
        std::map<std::array<int, 4>, int> smemo;
        std::function<int(int, int, int, int)> S = [&](int d, int i, int j, int e) -> int {
            std::array<int, 4> k = {d, i, j, e};
            auto it = smemo.find(k);
            if (it != smemo.end()) return it->second;
            int id;
            if (i + j + e == 0) {
                id = GP.add(R_IN, d);
                base_nodes.push_back(id);
            } else {
                id = GP.add(R_TEMP);
                std::vector<GTerm> t;
                if (i > 0) {
                    t.push_back({S(d, i - 1, j, e), SI_SID_V0 + d});
                    if (i - 1 > 0) t.push_back({S(d, i - 2, j, e), SI_SID_K0 + (i - 1) - 1});
                    if (j > 0) t.push_back({S(d, i - 1, j - 1, e), SI_SID_K0 + j - 1});
                    if (e > 0) t.push_back({S(d, i - 1, j, e - 1), SI_SID_K0 + e - 1});
                } else if (j > 0) {
                    t.push_back({S(d, 0, j - 1, e), SI_SID_V1 + d});
                    if (j - 1 > 0) t.push_back({S(d, 0, j - 2, e), SI_SID_K0 + (j - 1) - 1});
                    if (e > 0) t.push_back({S(d, 0, j - 1, e - 1), SI_SID_K0 + e - 1});
                } else {
                    t.push_back({S(d, 0, 0, e - 1), SI_SID_V2 + d});
                    if (e - 1 > 0) t.push_back({S(d, 0, 0, e - 2), SI_SID_K0 + (e - 1) - 1});
                }
                GP.nodes[id].terms = t;
            }
            smemo[k] = id;
            return id;
        };
        std::map<std::array<int, 3>, int> tmemo;
        std::function<int(int, int, int)> T = [&](int d, int i, int j) -> int {
            std::array<int, 3> k = {d, i, j};
            auto it = tmemo.find(k);
            if (it != tmemo.end()) return it->second;
            int id;
            if (i + j == 0) {
                id = GP.add(R_IN, 3 + d);
                base_nodes.push_back(id);
            } else {
                id = GP.add(R_TEMP);
                std::vector<GTerm> t;
                if (i > 0) {
                    t.push_back({T(d, i - 1, j), SI_SID_V0 + d});
                    if (i - 1 > 0) t.push_back({T(d, i - 2, j), SI_SID_K0 + (i - 1) - 1});
                    if (j > 0) t.push_back({T(d, i - 1, j - 1), SI_SID_K0 + j - 1});
                    t.push_back({S(d, i, j, 0), SI_SID_S0});
                    if (i - 1 > 0) t.push_back({S(d, i - 2, j, 0), SI_SID_K1 + (i - 1) - 1});
                } else {
                    t.push_back({T(d, 0, j - 1), SI_SID_V1 + d});
                    if (j - 1 > 0) t.push_back({T(d, 0, j - 2), SI_SID_K0 + (j - 1) - 1});
                    t.push_back({S(d, 0, j, 0), SI_SID_S0});
                    if (j - 1 > 0) t.push_back({S(d, 0, j - 2, 0), SI_SID_K2 + (j - 1) - 1});
                }
                GP.nodes[id].terms = t;
            }
            tmemo[k] = id;
            return id;
        };
        // operator components (defs/multipole.py:42-57)
        std::vector<I3> les;
        if (key == SI_KEY_OVLP || key == SI_KEY_KIN) les.push_back({0, 0, 0});
        else if (key == SI_KEY_DPM) les = comps_range(1, 1);
        else if (key == SI_KEY_QPM) les = comps_range(2, 2);
        else les = {I3{2, 0, 0}, I3{0, 2, 0}, I3{0, 0, 2}};
        x_size = ncomp * (int)ca.size() * (int)cb.size();
        int k = 0;
        for (int c = 0; c < ncomp; ++c)
            for (auto& a : ca)
                for (auto& b : cb) {
                    int id = GP.add(R_OUT, k++);  // X address (x_off = 0)
                    GNode& n = GP.nodes[id];
                    n.prod3 = true;
                    n.accum = true;
                    if (key == SI_KEY_KIN) {
                        int sx = S(0, a[0], b[0], 0), sy = S(1, a[1], b[1], 0), sz = S(2, a[2], b[2], 0);
                        int tx = T(0, a[0], b[0]), ty = T(1, a[1], b[1]), tz = T(2, a[2], b[2]);
                        GNode& nn = GP.nodes[id];
                        nn.terms = {{tx, 0}, {sy, 0}, {sz, 0}, {sx, 0}, {ty, 0}, {sz, 0}, {sx, 0}, {sy, 0}, {tz, 0}};
                    } else {
                        const I3& e = les[c];
                        int sx = S(0, a[0], b[0], e[0]), sy = S(1, a[1], b[1], e[1]), sz = S(2, a[2], b[2], e[2]);
                        GNode& nn = GP.nodes[id];
                        nn.terms = {{sx, 0}, {sy, 0}, {sz, 0}};
                    }
                }
    } else if (key == SI_KEY_COUL && !si_coul_use_hrr()) {
        // two-index VRR exactly as the reference's python path (defs/coulomb.py:35-92):
        // order x(bra), x(ket), y(bra), y(ket), z(bra), z(ket); no HRR.
        nlo = 0;
        nbase = L + 1;
        std::map<std::array<int, 7>, int> memo;
        std::function<int(const I3&, const I3&, int)> E = [&](const I3& a, const I3& b, int N) -> int {
            std::array<int, 7> k = {a[0], a[1], a[2], b[0], b[1], b[2], N};
            auto it = memo.find(k);
            if (it != memo.end()) return it->second;
            int id;
            if (sum3(a) + sum3(b) == 0) {
                if (N >= nbase) throw std::runtime_error("Boys order out of range");
                id = GP.add(R_IN, N);
                base_nodes.push_back(id);
            } else {
                int d = -1;
                bool bra = true;
                for (int q = 0; q < 3; ++q) {
                    if (a[q] > 0) { d = q; bra = true; break; }
                    if (b[q] > 0) { d = q; bra = false; break; }
                }
                I3 a2 = bra ? dec(a, d) : a;
                I3 b2 = bra ? b : dec(b, d);
                std::vector<GTerm> t;
                t.push_back({E(a2, b2, N), (bra ? SI_SID_V0 : SI_SID_V3) + d});  // PA | PB
                t.push_back({E(a2, b2, N + 1), SI_SID_V1 + d});                   // -PR
                if (a2[d] > 0) {
                    t.push_back({E(dec(a2, d), b2, N), SI_SID_K0 + a2[d] - 1});
                    t.push_back({E(dec(a2, d), b2, N + 1), SI_SID_K1 + a2[d] - 1});
                }
                if (b2[d] > 0) {
                    t.push_back({E(a2, dec(b2, d), N), SI_SID_K0 + b2[d] - 1});
                    t.push_back({E(a2, dec(b2, d), N + 1), SI_SID_K1 + b2[d] - 1});
                }
                id = GP.add(R_TEMP);
                GP.nodes[id].terms = t;
            }
            memo[k] = id;
            return id;
        };
        x_size = (int)ca.size() * (int)cb.size();
        for (auto& a : ca)
            for (auto& b : cb) targets.push_back(E(a, b, 0));
        finalize_targets(GP, targets, 0);
    } else if (key == SI_KEY_COUL || key == SI_KEY_3C2E) {
        const bool three = (key == SI_KEY_3C2E);
        const int lc = three ? Lc : 0;
        nlo = lc;
        nbase = L + 1;
        // VRR on a: (a, N)
        std::map<std::array<int, 4>, int> vmemo;
        std::function<int(const I3&, int)> V = [&](const I3& a, int N) -> int {
            std::array<int, 4> k = {a[0], a[1], a[2], N};
            auto it = vmemo.find(k);
            if (it != vmemo.end()) return it->second;
            int id;
            if (sum3(a) == 0) {
                if (N - nlo < 0 || N - nlo >= nbase) throw std::runtime_error("Boys order out of range");
                id = GP.add(R_IN, N - nlo);
                base_nodes.push_back(id);
            } else {
                int d = first_nz(a);
                I3 am = dec(a, d);
                std::vector<GTerm> t;
                t.push_back({V(am, N), SI_SID_V0 + d});      // PA
                t.push_back({V(am, N + 1), SI_SID_V1 + d});  // -(rho/p) PC  |  -PR
                int kq = a[d] - 1;
                if (kq > 0) {
                    I3 amm = dec(am, d);
                    t.push_back({V(amm, N), SI_SID_K0 + kq - 1});      // k/(2p)
                    t.push_back({V(amm, N + 1), SI_SID_K1 + kq - 1});  // -k rho/(2p^2) | -k/(2p)
                }
                id = GP.add(R_TEMP);
                GP.nodes[id].terms = t;
            }
            vmemo[k] = id;
            return id;
        };
        // truncated aux VRR on c: (a, c, N)   (defs/coulomb.py:293-308)
        std::map<std::array<int, 7>, int> dmemo;
        std::function<int(const I3&, const I3&, int)> Dn = [&](const I3& a, const I3& c, int N) -> int {
            if (sum3(c) == 0) return V(a, N);
            std::array<int, 7> k = {a[0], a[1], a[2], c[0], c[1], c[2], N};
            auto it = dmemo.find(k);
            if (it != dmemo.end()) return it->second;
            int d = first_nz(c);
            I3 cm = dec(c, d);
            std::vector<GTerm> t;
            t.push_back({Dn(a, cm, N + 1), SI_SID_V3 + d});  // (rho/c) PC_d
            if (a[d] > 0) t.push_back({Dn(dec(a, d), cm, N + 1), SI_SID_K2 + a[d] - 1});  // a_d/(2(p+c))
            int id = GP.add(R_TEMP);
            GP.nodes[id].terms = t;
            dmemo[k] = id;
            return id;
        };
        std::vector<I3> aps = comps_range(La, L);
        std::vector<I3> cc = comps_range(lc, lc);
        x_size = (int)aps.size() * (int)cc.size();
        for (auto& a : aps)
            for (auto& c : cc) targets.push_back(Dn(a, c, 0));
        finalize_targets(GP, targets, 0);
    } else if (key == SI_KEY_2C2E) {
        nlo = 0;
        nbase = L + 1;
        std::map<std::array<int, 7>, int> memo;
        std::function<int(const I3&, const I3&, int)> E = [&](const I3& a, const I3& b, int N) -> int {
            std::array<int, 7> k = {a[0], a[1], a[2], b[0], b[1], b[2], N};
            auto it = memo.find(k);
            if (it != memo.end()) return it->second;
            int id;
            if (sum3(a) + sum3(b) == 0) {
                if (N >= nbase) throw std::runtime_error("Boys order out of range");
                id = GP.add(R_IN, N);
                base_nodes.push_back(id);
            } else if (sum3(a) > 0) {
                int d = first_nz(a);
                I3 am = dec(a, d);
                std::vector<GTerm> t;
                t.push_back({E(am, b, N + 1), SI_SID_V0 + d});
                int kq = a[d] - 1;
                if (kq > 0) {
                    I3 amm = dec(am, d);
                    t.push_back({E(amm, b, N), SI_SID_K0 + kq - 1});
                    t.push_back({E(amm, b, N + 1), SI_SID_K1 + kq - 1});
                }
                if (b[d] > 0) t.push_back({E(am, dec(b, d), N + 1), SI_SID_K2 + b[d] - 1});
                id = GP.add(R_TEMP);
                GP.nodes[id].terms = t;
            } else {
                int d = first_nz(b);
                I3 bm = dec(b, d);
                std::vector<GTerm> t;
                t.push_back({E(a, bm, N + 1), SI_SID_V1 + d});
                int kq = b[d] - 1;
                if (kq > 0) {
                    I3 bmm = dec(bm, d);
                    t.push_back({E(a, bmm, N), SI_SID_K3 + kq - 1});
                    t.push_back({E(a, bmm, N + 1), SI_SID_K4 + kq - 1});
                }
                id = GP.add(R_TEMP);
                GP.nodes[id].terms = t;
            }
            memo[k] = id;
            return id;
        };
        x_size = (int)ca.size() * (int)cb.size();
        for (auto& a : ca)
            for (auto& b : cb) targets.push_back(E(a, b, 0));
        finalize_targets(GP, targets, 0);
    } else {
        throw std::runtime_error("unknown key");
    }

    // fix base addresses: base region right after X
    const int base_off = x_size;
    for (int id : base_nodes) GP.nodes[id].addr += base_off;
    const int a_off = base_off + nbase;
    int ptemp = B.emit(GP, a_off, false, 0, 0, 0, 0, 1, &B.P.n_fma_prim);
    D.n_prim_stages = (int)B.P.stages.size();

    // ---- contracted phase ---------------------------------------------------
    int n0 = 0, n1 = 0;
    int y_size = 0, nb_total = 1, out_size = 0;
    int in_off = 0, in_bs = 0, out_estride = 1, out_bs = 0;
    int y_off = a_off;
    if (key == SI_KEY_3C2E) {
        // C2S on c (batched over a'):  Y[m][a'] = sum_c M[m][c] X[a'][c]
        std::vector<I3> aps = comps_range(La, L);
        const int nap = (int)aps.size();
        const int ncc = si_ncart(Lc);
        int nm = 0;
        std::vector<double> Mc = final_transform(Lc, 1, normalize, nm);
        nb_total = nm;
        y_size = nm * nap;
        // group m's by term count
        std::map<int, std::vector<int>> by_nt;
        for (int m = 0; m < nm; ++m) {
            int nt = 0;
            for (int c = 0; c < ncc; ++c)
                if (Mc[(size_t)m * ncc + c] != 0.0) nt++;
            by_nt[nt].push_back(m);
        }
        bool first = true;
        for (auto& kv : by_nt) {
            SiStage st;
            std::memset(&st, 0, sizeof(st));
            st.hdr_off = (int)B.P.hdr.size();
            st.term_off = (int)B.P.terms.size();
            st.nelem = (int)kv.second.size();
            st.nterms = kv.first;
            st.batched = 0;
            st.nb_fixed = nap;  // own fixed batch count
            st.sbs = ncc;
            st.dbs = 1;
            st.flags = first ? SI_ST_SYNC : 0;
            st.magic = Builder::magic_for(st.nelem, st.nelem * nap);
            first = false;
            for (int m : kv.second) B.P.hdr.push_back((uint32_t)(y_off + m * nap));
            B.P.terms.resize(B.P.terms.size() + (size_t)st.nterms * st.nelem);
            for (int e = 0; e < st.nelem; ++e) {
                int m = kv.second[e];
                int k = 0;
                for (int c = 0; c < ncc; ++c) {
                    double v = Mc[(size_t)m * ncc + c];
                    if (v == 0.0) continue;
                    B.P.terms[st.term_off + (size_t)k * st.nelem + e] = (uint32_t)(0 + c) | ((uint32_t)B.const_sid(v) << 16);
                    k++;
                }
            }
            B.P.stages.push_back(st);
            B.P.n_fma_contr += (long long)st.nterms * st.nelem * nap;
        }
        in_off = y_off;
        in_bs = nap;
        out_estride = nm;
        out_bs = 1;
        D.out_batch_major = 0;
    } else if (is_1e) {
        nb_total = ncomp;
        in_off = 0;
        in_bs = (int)ca.size() * (int)cb.size();
        out_estride = 1;
        D.out_batch_major = 1;
    } else {
        nb_total = 1;
        in_off = 0;
        in_bs = 0;
        out_estride = 1;
        D.out_batch_major = 0;
    }
    D.chunk_begin = (int)B.P.stages.size();

    // output block size
    {
        int t0, t1;
        final_transform(La, sph, normalize, t0);
        final_transform(Lb, sph, normalize, t1);
        n0 = t0;
        n1 = t1;
    }
    out_size = n0 * n1 * nb_total;
    if (is_1e) out_bs = n0 * n1;
    const int out_off = y_off + y_size;

    Graph GC;
    ContrSpec cs;
    cs.La = La;
    cs.Lb = Lb;
    cs.hrr = ((key == SI_KEY_COUL && si_coul_use_hrr()) || key == SI_KEY_3C2E);
    cs.in_off = in_off;
    cs.out_off = out_off;
    cs.out_estride = out_estride;
    build_contr_graph(B, GC, cs, (key == SI_KEY_3C2E) ? 1 : sph, normalize, n0, n1);
    // first pass to learn the slot count (stages are discarded), then the real pass
    int ctemp_off;
    int chunk;
    {
        Builder probe = B;
        Graph Gp = GC;
        int slots = probe.emit(Gp, 0, true, in_bs, in_bs, out_bs, out_bs, 1, nullptr);
        if (key == SI_KEY_3C2E && slots > 0 && a_off / slots >= 1) {
            // X and the base values are dead after C2S(c): overlay the temporaries
            ctemp_off = 0;
            chunk = std::min(nb_total, a_off / slots);
        } else {
            ctemp_off = out_off + out_size;
            chunk = (key == SI_KEY_3C2E) ? 1 : nb_total;
        }
        int ctemp_size = slots * chunk;
        int w1 = a_off + ptemp;
        int w2 = out_off + out_size;
        int w3 = ctemp_off + ctemp_size;
        D.w_size = std::max(w1, std::max(w2, w3));
    }
    long long fc = 0;
    B.emit(GC, ctemp_off, true, in_bs, in_bs, out_bs, out_bs, nb_total, &fc);
    B.P.n_fma_contr += fc;

    D.nstages = (int)B.P.stages.size();
    D.nconst = (int)B.P.consts.size();
    D.nb_total = nb_total;
    D.chunk = chunk;
    D.x_off = 0;
    D.x_size = x_size;
    D.base_off = base_off;
    D.nbase = nbase;
    D.boys_nlo = nlo;
    D.out_off = out_off;
    D.out_n0 = n0;
    D.out_n1 = n1;
    if (D.w_size >= 65536) throw std::runtime_error("work array exceeds 16-bit addressing");
    return B.P;
}
