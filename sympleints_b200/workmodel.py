"""Algorithmic work model: flops and bytes per shell class.

The flop figures are the dependency-edge counts of the REFERENCE's own recurrence graphs
(sympleints.graphs.optimize.graphs_from_transforms on graphs/defs/int_{nucattr,2c2e,3c2e}.py;
one edge = one FMA = 2 flops), so ``roofline.achieved`` is quoted in the reference's own units
no matter how many operations our kernels actually execute:

    per primitive tuple : 2*(E_vrr + E_vrr2) + 12*n_boys + 40 + 2*S_contr
    per contracted tuple: 2*(E_c2s_c + E_hrr + E_c2s_b + E_c2s_a)
    n_boys = La+Lb+1, S_contr = ncart(Lc) * sum_{a=La}^{La+Lb} ncart(a)   (3c2e)
"""

_NUCATTR = """00: 0 0 1 1; 10: 6 0 3 3; 20: 30 0 6 8; 30: 74 0 10 16; 40: 150 0 15 28; 11: 30 18 9 9; 21: 86 36 18 24; 31: 180 60 30 48; 41: 324 90 45 84; 22: 176 158 48 40; 32: 326 258 80 80; 42: 530 382 120 140; 33: 552 644 160 112; 43: 852 936 240 196; 44: 1296 1794 420 252"""

_INT2C2E = """00: 0 0 1 1; 10: 3 0 3 3; 20: 15 0 6 8; 30: 37 0 10 16; 40: 83 0 15 28; 50: 151 0 21 46; 11: 3 12 9 9; 21: 18 27 18 24; 31: 49 48 30 48; 41: 102 75 45 84; 51: 191 108 63 138; 22: 45 129 48 40; 32: 101 231 80 80; 42: 202 363 120 140; 52: 355 525 168 230; 33: 115 619 160 112; 43: 225 984 240 196; 53: 407 1435 336 322; 44: 326 2185 420 252; 54: 572 3201 588 414; 55: 625 6180 966 506"""

_INT3C2E = """
000: 0 0 1 0 1 1; 001: 0 3 3 0 3 3; 002: 0 9 8 0 5 5; 003: 0 17 16 0 7 7; 004: 0 27 28 0 9 9; 005: 0 41 46 0 11 11
100: 6 0 3 0 3 3; 101: 6 12 9 0 9 9; 102: 6 39 24 0 15 15; 103: 6 75 48 0 21 21; 104: 6 120 84 0 27 27; 105: 6 184 138 0 33 33
200: 30 0 6 0 6 8; 201: 30 27 18 0 18 24; 202: 30 93 48 0 30 40; 203: 30 184 96 0 42 56; 204: 30 297 168 0 54 72; 205: 30 460 276 0 66 88
300: 74 0 10 0 10 16; 301: 78 48 30 0 30 48; 302: 78 171 80 0 50 80; 303: 78 347 160 0 70 112; 304: 78 567 280 0 90 144; 305: 78 886 460 0 110 176
400: 150 0 15 0 15 28; 401: 158 75 45 0 45 84; 402: 160 273 120 0 75 140; 403: 160 564 240 0 105 196; 404: 160 933 420 0 135 252; 405: 160 1471 690 0 165 308
110: 30 0 9 18 9 9; 111: 30 39 27 54 27 27; 112: 30 120 72 90 45 45; 113: 30 228 144 126 63 63; 114: 30 363 252 162 81 81; 115: 30 553 414 198 99 99
210: 78 0 16 36 18 24; 211: 78 75 48 108 54 72; 212: 78 237 128 180 90 120; 213: 78 456 256 252 126 168; 214: 78 729 448 324 162 216; 215: 78 1116 736 396 198 264
310: 158 0 25 60 30 48; 311: 160 123 75 180 90 144; 312: 160 396 200 300 150 240; 313: 160 772 400 420 210 336; 314: 160 1242 700 540 270 432; 315: 160 1911 1150 660 330 528
410: 290 0 36 90 45 84; 411: 296 183 108 270 135 252; 412: 298 597 288 450 225 420; 413: 298 1176 576 630 315 588; 414: 298 1905 1008 810 405 756; 415: 298 2947 1656 990 495 924
220: 160 0 31 158 48 40; 221: 160 150 93 474 144 120; 222: 160 462 248 790 240 200; 223: 160 881 496 1106 336 280; 224: 160 1404 868 1422 432 360; 225: 160 2141 1426 1738 528 440
320: 296 0 46 258 80 80; 321: 298 231 138 774 240 240; 322: 298 720 368 1290 400 400; 323: 298 1384 736 1806 560 560; 324: 298 2214 1288 2322 720 720; 325: 298 3387 2116 2838 880 880
420: 496 0 64 382 120 140; 421: 498 330 192 1146 360 420; 422: 500 1038 512 1910 600 700; 423: 500 2009 1024 2674 840 980; 424: 500 3228 1792 3438 1080 1260; 425: 500 4956 2944 4202 1320 1540
330: 498 0 74 698 160 112; 331: 500 378 222 2094 480 336; 332: 500 1161 592 3490 800 560; 333: 500 2217 1184 4886 1120 784; 334: 500 3537 2072 6282 1440 1008; 335: 500 5396 3404 7678 1760 1232
430: 806 0 100 1024 240 196; 431: 812 522 300 3072 720 588; 432: 814 1614 800 5120 1200 980; 433: 814 3097 1600 7168 1680 1372; 434: 814 4956 2800 9216 2160 1764; 435: 814 7580 4600 11264 2640 2156
440: 1234 0 145 2204 420 252; 441: 1234 765 435 6612 1260 756; 442: 1236 2343 1160 11020 2100 1260; 443: 1236 4474 2320 15428 2940 1764; 444: 1236 7143 4060 19836 3780 2268; 445: 1236 10901 6670 24244 4620 2772
"""


def _parse(text):
    out = {}
    for item in text.replace("\n", ";").split(";"):
        item = item.strip()
        if not item:
            continue
        k, v = item.split(":")
        out[tuple(int(c) for c in k.strip())] = tuple(int(x) for x in v.split())
    return out


NUCATTR = _parse(_NUCATTR)   # (La,Lb): E_vrr E_hrr E_c2s_b E_c2s_a
INT2C2E = _parse(_INT2C2E)   # (La,Lb): E_vrr E_vrr2 E_c2s_b E_c2s_a
INT3C2E = _parse(_INT3C2E)   # (La,Lb,Lc): E_vrr E_vrr2 E_c2s_c E_hrr E_c2s_b E_c2s_a


def ncart(l):
    return (l + 1) * (l + 2) // 2


def flops_3c2e(La, Lb, Lc):
    """(flops per primitive triple, flops per contracted triple), La >= Lb."""
    if La < Lb:
        La, Lb = Lb, La
    e = INT3C2E[(La, Lb, Lc)]
    s_contr = ncart(Lc) * sum(ncart(a) for a in range(La, La + Lb + 1))
    n_boys = La + Lb + 1
    return 2 * (e[0] + e[1]) + 12 * n_boys + 40 + 2 * s_contr, 2 * (e[2] + e[3] + e[4] + e[5])


def flops_coul(La, Lb):
    """(per primitive pair and charge, per shell pair)."""
    if La < Lb:
        La, Lb = Lb, La
    e = NUCATTR[(La, Lb)]
    s_contr = sum(ncart(a) for a in range(La, La + Lb + 1))
    return 2 * e[0] + 12 * (La + Lb + 1) + 40 + 2 * s_contr, 2 * (e[1] + e[2] + e[3])


def flops_2c2e(La, Lb):
    if La < Lb:
        La, Lb = Lb, La
    e = INT2C2E[(La, Lb)]
    return 2 * (e[0] + e[1]) + 12 * (La + Lb + 1) + 40 + 2 * ncart(La) * ncart(Lb), 2 * (e[2] + e[3])


def tensor_3c2e_work(orb_shells, aux_shells, aux_range=None):
    """Totals for the (a>=b)-unique 3c2e tensor: integrals, primitive triples, model flops, bytes."""
    import collections

    pair_cls = collections.Counter()
    pair_prims = collections.Counter()
    n = len(orb_shells)
    by_l = collections.defaultdict(list)
    for s in orb_shells:
        by_l[s.L].append(len(s.exps))
    # unique pairs i >= j, classified by (max L, min L); primitive-pair sums per class
    Ls = [s.L for s in orb_shells]
    Ks = [len(s.exps) for s in orb_shells]
    for i in range(n):
        for j in range(i + 1):
            key = (max(Ls[i], Ls[j]), min(Ls[i], Ls[j]))
            pair_cls[key] += 1
            pair_prims[key] += Ks[i] * Ks[j]
    sb, se = (0, len(aux_shells)) if aux_range is None else aux_range
    aux_cls = collections.Counter()
    aux_prims = collections.Counter()
    for s in aux_shells[sb:se]:
        aux_cls[s.L] += 1
        aux_prims[s.L] += len(s.exps)
    integrals = flops = prim_triples = triples = 0
    per_class = {}
    for (La, Lb), npair in pair_cls.items():
        for Lc, naux in aux_cls.items():
            fp, fc = flops_3c2e(La, Lb, Lc)
            ntrip = npair * naux
            nprim = pair_prims[(La, Lb)] * aux_prims[Lc]
            nint = ntrip * (2 * La + 1) * (2 * Lb + 1) * (2 * Lc + 1)
            f = nprim * fp + ntrip * fc
            per_class[(La, Lb, Lc)] = dict(triples=ntrip, prim_triples=nprim, integrals=nint, flops=f)
            integrals += nint
            flops += f
            prim_triples += nprim
            triples += ntrip
    return dict(integrals=integrals, flops=flops, prim_triples=prim_triples, triples=triples,
                bytes=8 * integrals, per_class=per_class)


# Measured time per model flop of the generated 3c2e kernels by auxiliary angular momentum Lc, relative to the
# flop-weighted mean (B200, benchmark system, class-major aux order; scripts/calibrate_lc_weights.py ->
# profiles/r02_lc_weights.json).  Low Lc runs below the mean rate (fewer integrals per primitive-level recurrence).
LC_TIME_WEIGHT = (1.313, 1.088, 0.978, 0.900, 0.918, 1.012)


def aux_shell_costs(orb_shells, aux_shells, weights=LC_TIME_WEIGHT):
    """Cost attributable to each auxiliary shell (for cost-balanced sharding): the model flops of all its shell
    triples, times the measured time-per-flop weight of its angular momentum (``weights=None``: plain flops)."""
    import collections

    pair_cls = collections.Counter()
    pair_prims = collections.Counter()
    Ls = [s.L for s in orb_shells]
    Ks = [len(s.exps) for s in orb_shells]
    for i in range(len(orb_shells)):
        for j in range(i + 1):
            key = (max(Ls[i], Ls[j]), min(Ls[i], Ls[j]))
            pair_cls[key] += 1
            pair_prims[key] += Ks[i] * Ks[j]
    cost_by = {}
    out = []
    for s in aux_shells:
        k = (s.L, len(s.exps))
        if k not in cost_by:
            c = 0
            for (La, Lb), npair in pair_cls.items():
                fp, fc = flops_3c2e(La, Lb, s.L)
                c += pair_prims[(La, Lb)] * len(s.exps) * fp + npair * fc
            cost_by[k] = c * (1.0 if weights is None else weights[s.L])
        out.append(cost_by[k])
    return out
