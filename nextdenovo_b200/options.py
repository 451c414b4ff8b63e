"""Option mirror of minimap2-nd for the overlap path.

ref: minimap2/options.c:4-64 (defaults), :80-97 (presets ava-ont / ava-pb), minimap2/main.c:102-111
(K/M/G suffixes), :177-205 (first pass: -x and --step), :209-363 (second pass).
"""
from dataclasses import dataclass, field

MM_F_NO_DIAG = 0x001
MM_F_NO_DUAL = 0x002
MM_F_CIGAR = 0x004
MM_F_NO_LJOIN = 0x400
MM_F_ALL_CHAINS = 0x800000
MM_F_NO_END_FLT = 0x10000000
MM_I_HPC = 0x1


@dataclass
class IdxOpt:                       # mm_idxopt_t, minimap2/minimap.h:102-106
    k: int = 15
    w: int = 10
    flag: int = 0
    bucket_bits: int = 14
    mini_batch_size: int = 50000000
    batch_size: int = 4000000000


@dataclass
class MapOpt:                       # subset of mm_mapopt_t, minimap2/minimap.h:108-170
    flag: int = 0
    seed: int = 11
    bw: int = 500
    max_gap: int = 5000
    max_chain_skip: int = 25
    max_chain_iter: int = 5000
    min_cnt: int = 3
    min_chain_score: int = 40
    mid_occ_frac: float = 2e-4
    min_mid_occ: int = 0
    mid_occ: int = 0
    max_occ: int = 0
    mini_batch_size: int = 500000000
    step: int = 0
    outraw: int = 0
    minlen: int = 500
    maxhan1: int = 5000
    maxhan2: int = 500
    dvt: int = 0
    n_threads: int = 3
    # base-level alignment (-c), minimap2/options.c:37-44
    a: int = 2
    b: int = 4
    q: int = 4
    e: int = 2
    q2: int = 24
    e2: int = 1
    sc_ambi: int = 1
    zdrop: int = 400
    zdrop_inv: int = 200
    end_bonus: int = -1
    min_dp_max: int = 80
    min_ksw_len: int = 200
    max_clip_ratio: float = 1.0
    # --step 2 (minimap2/options.c:50-62, main.c:194-198)
    minide: float = 0.05
    minmatch: int = 100
    mode: int = 2
    cn: int = 20
    kn: int = 17
    wn: int = 10
    outctn: int = 0
    extra: dict = field(default_factory=dict)


def parse_num(s: str) -> int:
    """ref: minimap2/main.c:102-111 (mm_parse_num)"""
    s = s.strip()
    mult = 1.0
    num = s
    if s and s[-1] in "gG":
        mult, num = 1e9, s[:-1]
    elif s and s[-1] in "mM":
        mult, num = 1e6, s[:-1]
    elif s and s[-1] in "kK":
        mult, num = 1e3, s[:-1]
    return int(float(num) * mult + .499)


def set_opt(preset, io: IdxOpt, mo: MapOpt) -> int:
    """ref: minimap2/options.c:80-97 (mm_set_opt); only the presets the overlap path uses."""
    if preset is None:
        d_io, d_mo = IdxOpt(), MapOpt()
        io.__dict__.update(d_io.__dict__)
        mo.__dict__.update(d_mo.__dict__)
        return 0
    if preset == "ava-ont":
        io.flag, io.k, io.w = 0, 15, 5
        mo.flag |= MM_F_ALL_CHAINS | MM_F_NO_DIAG | MM_F_NO_DUAL | MM_F_NO_LJOIN
        mo.min_chain_score, mo.max_gap, mo.max_chain_skip, mo.bw = 100, 10000, 25, 2000
    elif preset == "ava-pb":
        io.flag |= MM_I_HPC
        io.k, io.w = 19, 5
        mo.flag |= MM_F_ALL_CHAINS | MM_F_NO_DIAG | MM_F_NO_DUAL | MM_F_NO_LJOIN
        mo.min_chain_score, mo.max_gap, mo.max_chain_skip = 100, 10000, 25
    else:
        return -1
    return 0


def parse_args(argv):
    """Two-pass parse of the minimap2-nd options the NextDenovo pipeline passes for --step 1
    (ref: nextDenovo:445-464, lib/config_parser.py:191-217).  Returns (io, mo, out, positional)."""
    io, mo = IdxOpt(), MapOpt()
    takes = {"-x", "--step", "-k", "-w", "-t", "-I", "-K", "-o", "-r", "-g", "-n", "-m", "-f", "--minlen",
             "--maxhan1", "--maxhan2", "--seed", "--max-chain-skip", "--max-chain-iter", "--min-occ-floor", "-v",
             "-A", "-B", "-O", "-E", "-z", "-s", "--min-dp-score", "--end-bonus", "--max-clip-ratio",
             "--minide", "--minmatch", "--mode", "--kn", "--wn", "--cn"}
    toks, pos, i = [], [], 0
    while i < len(argv):
        a = argv[i]
        if a.startswith("--") and "=" in a:
            k, v = a.split("=", 1)
            toks.append((k, v))
        elif a in takes:
            if i + 1 >= len(argv):
                raise ValueError("[ERROR] missing option argument")
            toks.append((a, argv[i + 1]))
            i += 1
        elif a.startswith("-") and len(a) > 1:
            toks.append((a, None))
        else:
            pos.append(a)
        i += 1
    for k, v in toks:                                   # pass 1 (main.c:177-205)
        if k == "-x":
            if set_opt(v, io, mo) < 0:
                raise ValueError("[ERROR] unknown preset '%s'" % v)
        elif k == "--step":
            mo.step, mo.outraw = int(v), 1
            if mo.step == 1:
                mo.minlen = 500
            elif mo.step in (2, 3):
                mo.minlen = 2000
                if mo.step == 2:
                    mo.minide, mo.maxhan1, mo.maxhan2, mo.minmatch, mo.kn, mo.wn, mo.cn = 0.05, 5000, 500, 100, 17, 10, 20
            else:
                raise ValueError("[ERROR] unknown preset step '%s'" % v)
    out = None
    for k, v in toks:                                   # pass 2 (main.c:209-363)
        if k == "-w": io.w = int(v)
        elif k == "-k": io.k = int(v)
        elif k == "-H": io.flag |= MM_I_HPC
        elif k == "-r": mo.bw = parse_num(v)
        elif k == "-t": mo.n_threads = int(v)
        elif k == "-g": mo.max_gap = parse_num(v)
        elif k == "-n": mo.min_cnt = int(v)
        elif k == "-m": mo.min_chain_score = int(v)
        elif k == "-I": io.batch_size = parse_num(v)
        elif k == "-K": mo.mini_batch_size = parse_num(v)
        elif k == "-D": mo.flag |= MM_F_NO_DIAG
        elif k == "-P": mo.flag |= MM_F_ALL_CHAINS
        elif k == "-X": mo.flag |= MM_F_ALL_CHAINS | MM_F_NO_DIAG | MM_F_NO_DUAL | MM_F_NO_LJOIN
        elif k == "-c": mo.flag |= MM_F_CIGAR
        elif k == "-A": mo.a = int(v)
        elif k == "-B": mo.b = int(v)
        elif k in ("-s", "--min-dp-score"): mo.min_dp_max = int(v)
        elif k == "--end-bonus": mo.end_bonus = int(v)
        elif k == "--max-clip-ratio": mo.max_clip_ratio = float(v)
        elif k == "--no-end-flt": mo.flag |= MM_F_NO_END_FLT
        elif k == "-z":
            z = v.split(",")
            mo.zdrop = mo.zdrop_inv = int(z[0])
            if len(z) > 1: mo.zdrop_inv = int(z[1])
        elif k == "-O":
            z = v.split(",")
            mo.q = mo.q2 = int(z[0])
            if len(z) > 1: mo.q2 = int(z[1])
        elif k == "-E":
            z = v.split(",")
            mo.e = mo.e2 = int(z[0])
            if len(z) > 1: mo.e2 = int(z[1])
        elif k == "--minlen": mo.minlen = parse_num(v)
        elif k == "--maxhan1": mo.maxhan1 = parse_num(v)
        elif k == "--maxhan2": mo.maxhan2 = parse_num(v)
        elif k == "--dvt": mo.dvt = 1
        elif k == "--minide": mo.minide = float(v)
        elif k == "--minmatch": mo.minmatch = parse_num(v)
        elif k == "--mode": mo.mode = int(v)
        elif k == "--kn": mo.kn = int(v)
        elif k == "--wn": mo.wn = int(v)
        elif k == "--cn": mo.cn = int(v)
        elif k == "--outctn": mo.outctn = 1
        elif k == "--seed": mo.seed = int(v)
        elif k == "--max-chain-skip": mo.max_chain_skip = int(v)
        elif k == "--max-chain-iter": mo.max_chain_iter = int(v)
        elif k == "--min-occ-floor": mo.min_mid_occ = int(v)
        elif k == "--dual":                              # main.c:328-329: yes clears NO_DUAL
            if v in ("yes", "y"): mo.flag &= ~MM_F_NO_DUAL
            elif v in ("no", "n"): mo.flag |= MM_F_NO_DUAL
        elif k == "-f":
            x = float(v.split(",")[0])
            if x < 1.0: mo.mid_occ_frac, mo.mid_occ = x, 0
            else: mo.mid_occ = int(x + .499)
        elif k == "-o":
            if v != "-":
                out = v
                if mo.step: mo.outraw = 0
    return io, mo, out, pos
