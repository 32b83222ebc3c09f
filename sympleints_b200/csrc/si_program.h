// Host-side program builder: turns the recurrence definitions of one shell
// class into stage-machine programs (see si_machine.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "si_machine.h"

enum SiKey {
    SI_KEY_OVLP = 0,
    SI_KEY_KIN = 1,
    SI_KEY_DPM = 2,
    SI_KEY_QPM = 3,
    SI_KEY_DQPM = 4,
    SI_KEY_COUL = 5,
    SI_KEY_2C2E = 6,
    SI_KEY_3C2E = 7,
    SI_NKEYS = 8
};

// normalize: 0 none, 1 cgto (lmn factors), 2 pgto (lmn factors here; the radial
// part is folded into the coefficients by the caller, see si_abi.cu)
struct SiProgramHost {
    std::vector<uint32_t> hdr;
    std::vector<uint32_t> terms;
    std::vector<SiStage> stages;
    std::vector<double> consts;
    SiProg desc;  // pointer members are left null; owners patch them
    long long n_fma_prim = 0;   // linear terms executed per primitive tuple
    long long n_fma_contr = 0;  // linear terms executed per contracted tuple
};

int si_coul_use_hrr();
void si_coul_set_hrr(int v);
int si_ncart(int l);
int si_nsph(int l);
int si_ncomp(int key);
// Canonical Cartesian order (sympleints/helpers.py:42-49).
void si_cart_order(int l, std::vector<int>& lmn /* 3*ncart */);
// Real cart->sph matrix (sympleints/cart2sph.py:163-202), row-major (2l+1) x ncart(l).
std::vector<double> si_cart2sph(int l);
// 1/sqrt((2l-1)!!(2m-1)!!(2n-1)!!) per Cartesian component (sympleints/main.py:175-183).
std::vector<double> si_lmn_factors(int l);

SiProgramHost si_build_program(int key, int La, int Lb, int Lc, int sph, int normalize);
