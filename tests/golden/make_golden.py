#!/usr/bin/env python
"""Regenerates the golden fixtures from the UNMODIFIED reference (needs /root/reference and a built
oracle/_ref: `make -C oracle`).  Run from the repo root.  Committed outputs:

  reads_test.seed.2bit      seq_dump -f 1000 -s 1001 -b 10g -n 1 of test_data/reads_test.fa.gz (997 reads)
  G1_ava-ont.ovl            minimap2-nd --step 1 -x ava-ont              T T   (md5 9fce14e8...)
  G2_ava-ont_k17w10.ovl     minimap2-nd --step 1 -x ava-ont -k 17 -w 10  T T   (md5 270d357d...)
  G3_ava-pb.ovl             minimap2-nd --step 1 -x ava-pb               T T   (md5 7fb430f9...)
  sub120.2bit               first 120 reads of reads_test.seed.2bit
  S1_sub120_ava-ont.ovl, S3_sub120_ava-pb.ovl, S5_sub120_dual.ovl (sub120 x full with --dual=yes -I 20M: 2 index parts)
  S6_sub120_ava-ont_c.ovl, S7_sub120_ava-pb_c.ovl   the same inputs with -c (base-level alignment, regime B)
  S8_sub30_ava-ont_c.ovl    first 30 reads of sub120, -c -x ava-ont (CPU-sized case for the host alignment logic)
  G4_ava-ont_c.ovl          minimap2-nd --step 1 -c -x ava-ont T T (md5 4c7e8525..., ~6 min on 8 threads)
  S9_sub120_step2.ovl(.bl), G5_step2.ovl(.bl)   minimap2-nd -I 6G --step 2 -x ava-ont -k 17 -w 17 --minlen 2000 --maxhan1 5000 (sub120 / T)
"""
import os, subprocess, sys, tempfile, hashlib
sys.path.insert(0, os.getcwd())
import nextdenovo_b200 as nd

REF = "oracle/_ref"
G = "tests/golden"


def run(args, out):
    subprocess.check_call([REF + "/minimap2-nd", "--step", "1", "-t", "8"] + args + ["-o", out],
                          stderr=subprocess.DEVNULL)
    print(out, hashlib.md5(open(out, "rb").read()).hexdigest())


def main():
    with tempfile.TemporaryDirectory() as td:
        open(td + "/input.fofn", "w").write("/root/reference/test_data/reads_test.fa.gz\n")
        subprocess.check_call([REF + "/seq_dump", "-f", "1000", "-s", "1001", "-b", "10g", "-n", "1", "-d", td + "/db",
                               td + "/input.fofn"], stderr=subprocess.DEVNULL)
        os.replace(td + "/db/input.seed.001.2bit", G + "/reads_test.seed.2bit")
    T = G + "/reads_test.seed.2bit"
    run(["-x", "ava-ont", T, T], G + "/G1_ava-ont.ovl")
    run(["-x", "ava-ont", "-k", "17", "-w", "10", T, T], G + "/G2_ava-ont_k17w10.ovl")
    run(["-x", "ava-pb", T, T], G + "/G3_ava-pb.ovl")
    r = nd.read_2bit(T, max_reads=120)
    nd.write_2bit(G + "/sub120.2bit", r)
    S = G + "/sub120.2bit"
    run(["-x", "ava-ont", S, S], G + "/S1_sub120_ava-ont.ovl")
    run(["-x", "ava-pb", S, S], G + "/S3_sub120_ava-pb.ovl")
    run(["--dual=yes", "-I", "20M", "-x", "ava-ont", T, S], G + "/S5_sub120_dual.ovl")
    run(["-c", "-x", "ava-ont", S, S], G + "/S6_sub120_ava-ont_c.ovl")
    run(["-c", "-x", "ava-pb", S, S], G + "/S7_sub120_ava-pb_c.ovl")
    with tempfile.TemporaryDirectory() as td:
        nd.write_2bit(td + "/sub30.2bit", nd.read_2bit(S, max_reads=30))
        run(["-c", "-x", "ava-ont", td + "/sub30.2bit", td + "/sub30.2bit"], G + "/S8_sub30_ava-ont_c.ovl")
    a2 = ["-I", "6G", "--step", "2", "-t", "8", "-x", "ava-ont", "-k", "17", "-w", "17", "--minlen", "2000", "--maxhan1", "5000"]
    for src, dst in ((S, "S9_sub120_step2.ovl"), (T, "G5_step2.ovl")):      # --step 2 writes DST and DST.bl
        subprocess.check_call([REF + "/minimap2-nd"] + a2 + [src, src, "-o", G + "/" + dst], stderr=subprocess.DEVNULL)
        print(dst, hashlib.md5(open(G + "/" + dst, "rb").read()).hexdigest())
    if "--full" in sys.argv:
        run(["-c", "-x", "ava-ont", T, T], G + "/G4_ava-ont_c.ovl")


if __name__ == "__main__":
    main()
