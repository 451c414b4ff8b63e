import sys, numpy as np
sys.path.insert(0, '.')
import nextdenovo_b200 as nd
from nextdenovo_b200.formats import unpack_bases
reads = nd.read_2bit("tests/golden/sub120.2bit")
rng = np.random.default_rng(9); lut = np.frombuffer(b"ACGT", dtype=np.uint8)
with open(sys.argv[1], "wb") as fh:
    for r in range(40):
        w0, w1 = int(reads.word_off[r]), int(reads.word_off[r + 1])
        s = bytearray(lut[unpack_bases(reads.words[w0:w1], int(reads.length[r]))].tobytes())
        for _ in range(int(rng.integers(0, 5))):
            p, ln = int(rng.integers(0, len(s))), int(rng.integers(1, 40)); s[p:p + ln] = b"N" * len(s[p:p + ln])
        if r % 4 == 0 and len(sys.argv) < 3: s = bytearray(bytes(s).lower())
        fh.write(b">%d\n" % int(reads.ids[r]) + bytes(s) + b"\n")
