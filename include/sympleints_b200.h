/*
 * sympleints_b200 -- C ABI of the B200 evaluation engine for sympleints' shell integrals.
 *
 * This is the drop-in boundary for the hot path of eljost/sympleints: what the
 * reference's generated per-class functions compute when called on batches of
 * shells.  Plain pointers and sizes only; no torch / C++ types.  All functions
 * return an int status (SI_OK == 0) and never throw across the ABI; use
 * si_strerror() for a message.  The library has no CPU fallback: every entry
 * point that computes integrals runs CUDA kernels on the device selected in
 * si_init() and returns SI_ERR_CUDA if that is not possible.
 *
 * Reference interfaces replaced (file:line in the reference repository):
 *   - generated callables  <name>_<La><Lb>[<Lc>](ax, da, A, bx, db, B[, cx, dc, C], R, result)
 *       sympleints/templates/py_func.tpl:1-16, argument order sympleints/Functions.py:67-87
 *       -> si_eval_block()
 *   - the "equivalent" La<Lb wrappers + resort_ba_ab / resort_bac_abc
 *       sympleints/templates/py_equi_func.tpl, py_module.tpl:23-38,
 *       graphs/templates/int_3c2e_mod.tpl:48-66 -> handled inside every entry point
 *   - the shell-pair driver intor_2c(shells, func_dict, ncomponents, spherical, R=...)
 *       sympleints/testing/intor.py:18-76 -> si_int1e(), si_coul(), si_int2c2e()
 *   - the 3-index loop nest of the graph/Fortran path (no driver exists in the reference)
 *       sympleints/graphs/templates/int_3c2e_func.tpl:30-80 -> si_int3c2e()
 *   - Boys function table of get_boys_func(nmax)  sympleints/testing/boys.py:44-142
 *       -> si_init(boys_table, ...) / si_boys_table()
 *   - Fortran module ABI  subroutine <name>(La, Lb, axs, das, A, bxs, dbs, B, R, result) and
 *     the mandatory <name>_init()  sympleints/templates/fortran_module.tpl:13-27 -> si_init()
 */
#ifndef SYMPLEINTS_B200_H
#define SYMPLEINTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define SI_OK 0
#define SI_ERR_ARG 1        /* null pointer / inconsistent sizes */
#define SI_ERR_LMAX 2       /* angular momentum exceeds lmax / lauxmax of the context */
#define SI_ERR_CUDA 3       /* CUDA runtime error (no device, launch failure, ...) */
#define SI_ERR_KEY 4        /* unknown integral key or unsupported option combination */
#define SI_ERR_BOYS 5       /* Boys order exceeds the table (IndexError in the reference) */
#define SI_ERR_INTERNAL 6

/* integral keys (sympleints/main.py:96-108, names of the generated modules in parentheses) */
#define SI_OVLP 0  /* ovlp3d */
#define SI_KIN 1   /* kinetic3d */
#define SI_DPM 2   /* dipole3d, 3 components x,y,z */
#define SI_QPM 3   /* quadrupole3d, 6 components xx,xy,xz,yy,yz,zz */
#define SI_DQPM 4  /* diag_quadrupole3d, 3 components xx,yy,zz */
#define SI_COUL 5  /* coulomb3d */
#define SI_2C2E 6  /* int2c2e3d */
#define SI_3C2E 7  /* int3c2e3d_sph */
#define SI_GTO 8        /* {cart,sph}_gto3d: f(ax, da, A, R, result), values of the shell's functions at the point R
                           (sympleints/main.py:520-549, defs/gto.py); si_eval_block ignores the b arguments */
#define SI_MULTI_SPH 10 /* multipole3d_sph: 9 components R_lm, l <= 2, m = -l..l, about the centre P of every primitive
                           pair (sympleints/defs/multipole.py:60-138, main.py:752-800); spherical + normalised output only */
#define SI_PREFACTOR 9  /* prefactors: da db rP^(a+b), normalised, C2S on both indices (main.py:481-518,
                           defs/overlap.py:16-34); rP is passed as R (a module-level name in the reference's code) */

/* --normalize (sympleints/main.py:109-114) */
#define SI_NORM_NONE 0
#define SI_NORM_CGTO 1
#define SI_NORM_PGTO 2

/* layouts of the 3-index tensor */
#define SI_LAYOUT_FULL 0    /* (nbf, nbf, naux_local), C order, both triangles filled */
#define SI_LAYOUT_PACKED 1  /* (nrows_packed, naux_local): rows = shell pairs i>=j, each a dense (nf_i, nf_j) block */

typedef struct si_context si_context;
typedef struct si_basis si_basis;

const char* si_strerror(int status);
const char* si_last_error(void); /* detail of the last failure on this thread */

/*
 * Create a context on CUDA device `device`.  lmax/lauxmax bound the classes that may be
 * requested (--lmax / --lauxmax).  boys_table may be NULL: the library then builds the table
 * itself with the reference algorithm (top row n = nmax+6 on x = 0,0.1,...,30, downward
 * recursion on the grid); otherwise it must be the row-major (nrows, ncols) table of
 * get_boys_func(nmax = nrows-7) with ncols == 301.
 */
int si_init(int device, int lmax, int lauxmax, const double* boys_table, int nrows, int ncols,
            si_context** ctx);
int si_finalize(si_context* ctx);
/* Copy out the Boys table in use (host buffer of nrows*ncols doubles); sizes via pointers. */
int si_boys_table(const si_context* ctx, double* table, int* nrows, int* ncols);
/* Evaluate F_n(x) for n_x arguments ON THE DEVICE (host in / host out) -- parity check of the Boys kernel. */
int si_boys_eval(si_context* ctx, int n, const double* x, size_t n_x, double* out);
/* The same for the evaluation variants the generated kernels use on the hot path (same table, cell and six
 * Taylor terms; only the operation order differs): mode 0 = reference summation order (== si_boys_eval),
 * 1 = Horner/FMA form + binary-power asymptote (3c2e / 2c2e kernels), 2 = shared-rsqrt asymptote (coul kernels). */
int si_boys_eval_mode(si_context* ctx, int mode, int n, const double* x, size_t n_x, double* out);

/*
 * Shell set (struct of arrays, host pointers; copied to the device).  coefs are the
 * contraction coefficients exactly as the reference's generated functions receive them
 * (for --normalize cgto: already multiplied by the radial norms, sympleints/testing/shells.py:17-18).
 * exps/coefs are concatenated over shells in shell order; nprim[i] entries each.
 */
int si_basis_create(si_context* ctx, int nshell, const int* L, const int* nprim, const double* centers,
                    const double* exps, const double* coefs, si_basis** basis);
int si_basis_destroy(si_basis* basis);
int si_basis_nbf(const si_basis* basis, int sph);     /* number of basis functions */
int si_basis_nshell(const si_basis* basis);
/* number of rows of the packed 3-index layout: sum over shell pairs i>=j of nf_i*nf_j */
long long si_basis_packed_rows(const si_basis* basis);
/*
 * Restrict the shell pairs the device entry points work on to (i, j <= i), shell_begin <= i < shell_end, optionally
 * without the diagonal pairs: 2-index matrices then receive only those blocks (and their mirror images), the packed
 * 3-index layout holds only those rows (si_basis_selected_rows).  The per-class micro-benchmark uses it to build pure
 * same-class batches (the reference's benchmark.py:65-142 times one class at a time); (0, nshell, 0) restores all pairs.
 */
int si_basis_select_rows(si_basis* basis, int shell_begin, int shell_end, int skip_diagonal);
long long si_basis_selected_rows(const si_basis* basis);

/*
 * One-electron matrices over a shell set: out[(comp, mu, nu)], C order, shape (ncomp, nbf, nbf),
 * both triangles filled (what intor_2c returns).  key in {SI_OVLP, SI_KIN, SI_DPM, SI_QPM, SI_DQPM, SI_MULTI_SPH}.
 * R: multipole origin (3 doubles, host).  out is a DEVICE pointer; stream is a cudaStream_t (may be NULL).
 */
int si_int1e(si_context* ctx, const si_basis* basis, int key, int sph, int normalize, const double* R,
             double* out_dev, void* stream);
/* The whole one-electron set in one call (one kernel launch for sph + cgto|pgto): out_dev = (11, nbf, nbf):
 * component 0 ovlp, 1 kin, 2-4 dpm x,y,z, 5-10 qpm xx,xy,xz,yy,yz,zz. */
int si_int1e_set(si_context* ctx, const si_basis* basis, int sph, int normalize, const double* R, double* out_dev,
                 void* stream);
/*
 * Nuclear attraction / external potential: out[mu,nu] = sum_C q[C] * (mu| 1/|r-R_C| |nu).
 * The generated coulomb3d functions return the integral for a unit positive charge
 * (sympleints/defs/coulomb.py:79-80); callers pass q = -Z for nuclei.  charges: ncharge*3 host doubles.
 */
int si_coul(si_context* ctx, const si_basis* basis, int sph, int normalize, int ncharge,
            const double* charges_xyz, const double* q, double* out_dev, void* stream);
/* 2-centre-2-electron metric over an auxiliary shell set, (naux, naux). */
int si_int2c2e(si_context* ctx, const si_basis* aux, int sph, int normalize, double* out_dev, void* stream);
/*
 * 3-centre-2-electron tensor (ab|c), always spherical (sympleints/main.py:934-940), for the
 * auxiliary shells [aux_shell_begin, aux_shell_end) -- the unit of multi-GPU sharding.
 * naux_local = number of auxiliary functions of that shell range.
 */
int si_int3c2e(si_context* ctx, const si_basis* orb, const si_basis* aux, int aux_shell_begin,
               int aux_shell_end, int normalize, int layout, double* out_dev, size_t out_len, void* stream);

/*
 * Shell functions on a grid (the batched form of the `gto` key): out[(mu, g)] = chi_mu(r_g), shape (nbf, npoints),
 * C order; points_xyz: npoints*3 host doubles.  si_gto: DEVICE output, si_gto_h: host output.
 */
int si_gto(si_context* ctx, const si_basis* basis, int sph, int normalize, long long npoints, const double* points_xyz,
           double* out_dev, void* stream);
int si_gto_h(si_context* ctx, const si_basis* basis, int sph, int normalize, long long npoints,
             const double* points_xyz, double* out_host);

/* Host-buffer variants: allocate device scratch, run, copy the result back (synchronous). */
int si_int1e_h(si_context* ctx, const si_basis* basis, int key, int sph, int normalize, const double* R,
               double* out_host);
int si_coul_h(si_context* ctx, const si_basis* basis, int sph, int normalize, int ncharge,
              const double* charges_xyz, const double* q, double* out_host);
int si_int2c2e_h(si_context* ctx, const si_basis* aux, int sph, int normalize, double* out_host);
int si_int3c2e_h(si_context* ctx, const si_basis* orb, const si_basis* aux, int aux_shell_begin,
                 int aux_shell_end, int normalize, int layout, double* out_host, size_t out_len);

/*
 * Schwarz integrals (section 8f-2; sympleints/graphs/defs/int_4c2e.py, l_iters.py:23-28 `schwarz_iter`): D_dev (nbf, nbf)
 * receives D[mu, nu] = (mu nu|mu nu) for the spherical functions of the set, Q_dev (nshell, nshell; may be NULL) the
 * shell-pair factors Q[a, b] = sqrt(max_{mu in a, nu in b} D[mu, nu]), so that |(ab|P)| <= Q[a, b] sqrt((P|P)).
 * Needs Boys orders up to 4 lmax (SI_ERR_BOYS if the caller's table is smaller).
 */
int si_schwarz(si_context* ctx, const si_basis* basis, int normalize, double* D_dev, double* Q_dev, void* stream);

/*
 * Opt-in primitive pre-screening of the 3c2e kernels (eps = 0 switches it off; that is the default, the parity mode and
 * the benchmark mode -- the reference's own estimate, graphs/templates/int_3c2e_func.tpl:46-54, is a no-op `continue`).
 * A primitive triple is skipped when prefactor^2 * min(1, pi / (4 T)) <= eps^2 for all tuples of a warp, i.e. when the
 * bound of its (ss|s)-type base integral [2 pi^2.5 / (p c sqrt(p+c)) K_ab d_a d_b d_c F_0(T)] is below eps.
 */
int si_set_screening(si_context* ctx, double eps);

/* Timing of the last si_int3c2e_h call on this context, seconds: out[0] total, [1] host shell-batcher (0 when the
 * row-chunk pair sets were cached), [2] staging allocation (0 when large enough), [3] host time to enqueue the chunks,
 * [4] device time of the compute chunks, [5] wait for D2H copies after the last kernel, [6] chunks, [7] bytes per chunk. */
int si_host_breakdown(const si_context* ctx, double* out8);

/*
 * One shell block with the reference's generated-function signature
 *   f(ax, da, A, bx, db, B, [cx, dc, C,] R, result)
 * (host pointers; Ka/Kb/Kc primitives; result is caller-owned, flat, C order (comp, a, b[, c]),
 * overwritten).  Any La/Lb order is accepted (the La<Lb wrappers of the reference are folded in).
 * For SI_COUL R is the position of a unit positive charge; for the multipole keys the origin;
 * otherwise ignored (may be NULL).  Runs on the device.
 */
int si_eval_block(si_context* ctx, int key, int La, int Lb, int Lc, int sph, int normalize, int Ka,
                  const double* ax, const double* da, const double* A, int Kb, const double* bx,
                  const double* db, const double* B, int Kc, const double* cx, const double* dc,
                  const double* C, const double* R, double* result);

/* Measured FP64 FMA peak of the device (TFLOP/s, best of several launches over ~`seconds`):
 * the roofline denominator for the Boys-type kernels. */
int si_fp64_peak(si_context* ctx, double seconds, double* tflops);
/* Bookkeeping for benchmarks: kernels launched by this context since creation. */
long long si_launch_count(const si_context* ctx);
/*
 * Work model of one class: terms (FMA) executed per primitive tuple and per contracted tuple,
 * shared-memory doubles per tuple in flight, look-up-table words.  out[0..3].
 */
int si_class_stats(si_context* ctx, int key, int La, int Lb, int Lc, int sph, int normalize, long long* out);

#ifdef __cplusplus
}
#endif
#endif
