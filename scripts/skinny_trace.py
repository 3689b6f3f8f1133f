"""Reads the per-CTA phase stamps written by gemm_skinny_kernel under B2_SKINNY_TRACE=<file> (csrc/gemm_skinny.cu) and prints,
for the launches of the LAST decode step in the file, where each launch's time goes:

    span      last exit - first entry of the launch              overlap   first entry of the NEXT launch - last exit of this one
    head      entry -> upstream grid complete (pdl_wait returns)  stream    upstream complete -> last MMA issued
    drain     last MMA issued -> last tile's accumulator ready    publish   accumulator ready -> partial written / finaliser elected
    tail      publish -> exit (finalisers: re-read partials, epilogue, stores); fx_wait / fx_epi / fx_exit split a staged fix-up into
              partials landed in shared memory / sums + stores issued / exit

    B2_SKINNY_TRACE=gpurun_out/sk_trace.txt python scripts/decode_ab.py --batches 32 --variants REPS=2 ...
    python scripts/skinny_trace.py gpurun_out/sk_trace.txt [launches_per_step]
"""
import statistics
import sys


def main():
    path = sys.argv[1]
    per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 129
    launches = []
    cur = None
    for line in open(path):
        f = line.split()
        if f[0] == "launch":
            cur = dict(idx=int(f[1]), N=int(f[3]), K=int(f[5]), B=int(f[7]), grid=int(f[9]), rows=[])
            launches.append(cur)
        else:
            cur["rows"].append([int(x) for x in f[1:]])
    step = launches[-per_step:]
    med = statistics.median
    print(f"{len(launches)} launches in the file; analysing the last {len(step)}")
    print("  # N      K      span  entry_spread head(med/max) stream(med/max) drain(med) publish(med) tail_final(med/max) tail_other(med) finalisers next_overlap  [us]")
    agg = {}
    for i, L in enumerate(step):
        rows = [r for r in L["rows"] if r[0] != 0 and r[7] != 0]
        if not rows:
            continue
        t0 = min(r[0] for r in rows)
        t1 = max(r[7] for r in rows)
        us = lambda ns: ns / 1e3
        head = [r[2] - r[0] for r in rows]
        stream = [r[4] - r[2] for r in rows]
        drain = [r[5] - r[4] for r in rows]
        publish = [(r[6] - r[5]) if r[6] else 0 for r in rows]
        tail_f = [r[7] - max(r[6], r[5]) for r in rows if r[8] > 0]
        tail_o = [r[7] - max(r[6], r[5]) for r in rows if r[8] == 0]
        staged = [r for r in rows if len(r) > 10 and r[8] > 0 and r[9] and r[10]]
        fx_wait = [r[9] - max(r[6], r[5]) for r in staged]
        fx_epi = [r[10] - r[9] for r in staged]
        fx_exit = [r[7] - r[10] for r in staged]
        nxt = None
        if i + 1 < len(step):
            nr = [r for r in step[i + 1]["rows"] if r[0] != 0]
            if nr:
                nxt = min(r[0] for r in nr) - t1
        rec = dict(span=us(t1 - t0), spread=us(max(r[0] for r in rows) - t0), head=us(med(head)), head_max=us(max(head)),
                   stream=us(med(stream)), stream_max=us(max(stream)), drain=us(med(drain)), publish=us(med(publish)),
                   tail_f=us(med(tail_f)) if tail_f else 0.0, tail_f_max=us(max(tail_f)) if tail_f else 0.0,
                   tail_o=us(med(tail_o)) if tail_o else 0.0, nfin=len(tail_f), fx_wait=us(med(fx_wait)) if fx_wait else 0.0,
                   fx_epi=us(med(fx_epi)) if fx_epi else 0.0, fx_exit=us(med(fx_exit)) if fx_exit else 0.0, overlap=us(nxt) if nxt is not None else float("nan"))
        agg.setdefault((L["N"], L["K"]), []).append(rec)
        if 40 <= i < 48 or i >= len(step) - 1:
            print(f"{i:3d} {L['N']:6d} {L['K']:6d} {rec['span']:7.1f} {rec['spread']:9.1f}   {rec['head']:5.1f}/{rec['head_max']:5.1f}   "
                  f"{rec['stream']:6.1f}/{rec['stream_max']:6.1f}   {rec['drain']:6.1f}   {rec['publish']:6.1f}      "
                  f"{rec['tail_f']:5.1f}/{rec['tail_f_max']:5.1f}        {rec['tail_o']:5.1f}       {rec['nfin']:4d}     {rec['overlap']:7.1f}")
    print("\nmedians over the step, per GEMM shape:")
    for (N, K), recs in agg.items():
        line = f"N={N:6d} K={K:6d} n={len(recs):3d} " + " ".join(f"{k}={med([r[k] for r in recs]):.1f}" for k in
                                                              ("span", "spread", "head", "stream", "drain", "publish", "tail_f", "tail_f_max", "tail_o", "fx_wait", "fx_epi", "fx_exit", "overlap"))
        print(line)
    tot = sum(r["span"] for recs in agg.values() for r in recs)
    print(f"sum of spans: {tot / 1e3:.3f} ms")


if __name__ == "__main__":
    main()
