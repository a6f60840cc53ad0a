"""Debug helper (not a test): per-round GPU vs oracle comparison on a small data set."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from util import *
import goldpolish_b200 as gp

d = dataset(genome_len=int(sys.argv[1]) if len(sys.argv) > 1 else 60000)
pl = plan(d, 1)
ref = oracle_build(d, pl)
nb = len(pl.batch_entry_off) - 1
bad = 0
for ki, k in enumerate(KS):
    ctx = gp.Context(ks=(k,))
    pay = np.stack([ref[b].bfs[ki] for b in range(nb)])[:, None, :]
    ctx.load_filters(pay)
    # inputs of this round = oracle outputs of the previous rounds
    ins = []
    for c in range(d.n_contigs):
        cur = d.contig(c)
        for kj in range(ki):
            if cur is None: break
            cur, _ = ol.ntedit_contig(cur, ref[pl.contig_batch[c]].bfs[kj], KS[kj])
        ins.append(cur if cur is not None else b"")
    seq = np.frombuffer(b"".join(ins), dtype=np.uint8)
    off = np.cumsum([0] + [len(x) for x in ins]).astype(np.uint64)
    out, ooff, dropped = ctx.polish(seq, off, pl.contig_batch)
    st = ctx.stats()
    tot = {}
    for c in range(d.n_contigs):
        want, s = ol.ntedit_contig(ins[c], ref[pl.contig_batch[c]].bfs[ki], k)
        for kk, v in s.items(): tot[kk] = tot.get(kk, 0) + v
        got = out[int(ooff[c]):int(ooff[c + 1])].tobytes()
        if want is None:
            continue
        if got != want:
            bad += 1
            p = next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), min(len(got), len(want)))
            print(f"k={k} contig {c} len {len(ins[c])}: got {len(got)} want {len(want)} first diff at {p}")
            print("   in  ", ins[c][max(0, p - 40):p + 40])
            print("   got ", got[max(0, p - 40):p + 40])
            print("   want", want[max(0, p - 40):p + 40])
    print(f"k={k}: gpu stats", {x: st[x] for x in ("triggers", "edits", "masked", "rollbacks")}, "oracle", tot)
    ctx.close()
print("mismatching (contig, round) pairs:", bad)
