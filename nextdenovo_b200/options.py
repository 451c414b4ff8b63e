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


# Option grammar of the reference (ref: minimap2/main.c:31-100 long_options, :162 opt_str).  name -> (arity, canonical short form)
SHORT_SPEC = "2aSDw:k:K:t:r:f:Vv:g:G:I:d:XT:s:x:Hcp:M:n:z:A:B:O:E:m:N:Qu:R:hF:LC:yYPo:"
LONG_OPTS = {n: (a, c) for n, a, c in [
    ("bucket-bits", 1, None), ("mb-size", 1, "-K"), ("seed", 1, None), ("no-kalloc", 0, None), ("print-qname", 0, None), ("no-self", 0, "-D"),
    ("print-seeds", 0, None), ("max-chain-skip", 1, None), ("min-dp-len", 1, None), ("print-aln-seq", 0, None), ("splice", 0, None),
    ("cost-non-gt-ag", 1, "-C"), ("no-long-join", 0, None), ("sr", 0, None), ("frag", 1, None), ("secondary", 1, None), ("cs", 2, None),
    ("end-bonus", 1, None), ("no-pairing", 0, None), ("splice-flank", 1, None), ("idx-no-seq", 0, None), ("end-seed-pen", 1, None),
    ("for-only", 0, None), ("rev-only", 0, None), ("heap-sort", 1, None), ("all-chain", 0, "-P"), ("dual", 1, None), ("max-clip-ratio", 1, None),
    ("min-occ-floor", 1, None), ("MD", 0, None), ("lj-min-ratio", 1, None), ("score-N", 1, None), ("eqx", 0, None), ("paf-no-hit", 0, None),
    ("split-prefix", 1, None), ("no-end-flt", 0, None), ("hard-mask-level", 0, None), ("cap-sw-mem", 1, None), ("max-qlen", 1, None),
    ("max-chain-iter", 1, None), ("junc-bed", 1, None), ("junc-bonus", 1, None), ("sam-hit-only", 0, None), ("help", 0, "-h"),
    ("max-intron-len", 1, "-G"), ("version", 0, "-V"), ("min-count", 1, "-n"), ("min-chain-score", 1, "-m"), ("mask-level", 1, "-M"),
    ("min-dp-score", 1, "-s"), ("sam", 0, "-a"),
    ("step", 1, None), ("minlen", 1, None), ("minide", 1, None), ("maxhan1", 1, None), ("maxhan2", 1, None), ("outraw", 0, None),
    ("minmatch", 1, None), ("mode", 1, None), ("kn", 1, None), ("wn", 1, None), ("cn", 1, None), ("outctn", 0, None), ("df", 1, None), ("dvt", 0, None),
    ("device", 1, None), ("gpus", 1, None)]}
# options that change results and are not implemented: refused by name, never approximated
NOT_BUILT = {"-a", "-S", "--cs", "--MD", "--eqx", "-Y", "-L", "-Q", "-y", "-R", "--splice", "--sr", "--frag", "--no-pairing", "--splice-flank", "-u", "-C",
             "-G", "-F", "--heap-sort", "--for-only", "--rev-only", "--end-seed-pen", "--lj-min-ratio", "--split-prefix", "--junc-bed", "--junc-bonus",
             "--sam-hit-only", "--paf-no-hit", "--secondary", "-T", "-d", "--bucket-bits", "--idx-no-seq", "--cap-sw-mem", "--max-qlen", "--print-seeds",
             "--print-aln-seq", "--print-qname", "--df", "--outraw", "--hard-mask-level"}


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
    elif preset == "ava-hifi":                           # ref: minimap2/options.c:98-109 (long k-mers: sketch.c:283-356)
        io.flag |= MM_I_HPC
        io.k, io.w = 51, 51
        mo.a, mo.b, mo.q, mo.q2, mo.e, mo.e2 = 1, 4, 6, 26, 2, 1
        mo.mid_occ_frac = 1e-4
        mo.flag |= MM_F_ALL_CHAINS | MM_F_NO_DIAG | MM_F_NO_DUAL | MM_F_NO_LJOIN
        mo.min_chain_score, mo.max_gap, mo.max_chain_skip = 100, 10000, 25
    else:
        return -1
    return 0


def parse_args(argv):
    """Two-pass parse of the minimap2-nd options the NextDenovo pipeline passes for --step 1
    (ref: nextDenovo:445-464, lib/config_parser.py:191-217).  Returns (io, mo, out, positional)."""
    io, mo = IdxOpt(), MapOpt()
    toks, pos, i = [], [], 0
    while i < len(argv):
        a = argv[i]
        if a == "-" or not a.startswith("-"):
            pos.append(a)
        elif a == "--":
            pos += argv[i + 1:]
            break
        elif a.startswith("--"):
            name, eq, val = a[2:].partition("=")
            if name not in LONG_OPTS:
                raise ValueError('[ERROR] unknown option in "%s"' % a)
            has_arg, canon = LONG_OPTS[name]
            if not eq:
                val = None
                if has_arg == 1:
                    if i + 1 >= len(argv):
                        raise ValueError("[ERROR] missing option argument")
                    val = argv[i + 1]
                    i += 1
            toks.append((canon or "--" + name, val))
        else:
            j = 1
            while j < len(a):
                c = a[j]
                at = SHORT_SPEC.find(c)
                if at < 0 or c == ":":
                    raise ValueError('[ERROR] unknown option in "%s"' % a)
                if SHORT_SPEC[at + 1:at + 2] == ":":
                    if j + 1 < len(a):
                        val = a[j + 1:]
                    else:
                        if i + 1 >= len(argv):
                            raise ValueError("[ERROR] missing option argument")
                        val = argv[i + 1]
                        i += 1
                    toks.append(("-" + c, val))
                    break
                toks.append(("-" + c, None))
                j += 1
        i += 1
    for k, _ in toks:
        if k in NOT_BUILT:
            raise ValueError("[ERROR] option '%s' is not built in this GPU build" % k)
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
        elif k == "-s": mo.min_dp_max = int(v)
        elif k == "--min-dp-len": mo.min_ksw_len = int(v)
        elif k == "--score-N": mo.sc_ambi = int(v)
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
