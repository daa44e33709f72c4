"""Summarise a tools/tc_timeline.py dump: per layer-half of the MMA warp: span, cycles waiting on the epilogue (2000) and on
weight stages (3000), remainder = issue + back-pressure; and the epilogue's per-quarter latencies."""
import sys
rows = [l.split() for l in open(sys.argv[1]) if l.strip() and l.split()[0].lstrip("-").isdigit()]
mma = [(int(r[0]), int(r[2])) for r in rows if r[1] == "MMA"]
out = []; cur = None
for t, c in mma:
    if 1000 <= c < 2000: cur = [c - 1000, t, None, 0, 0]
    elif 5000 <= c < 6000 and cur: cur[2] = t
    elif c == 2000 and cur: cur[3] = t
    elif c == 3000 and cur:
        cur[4] = t; out.append(cur); cur = None
print("layer.half  start   span  wait_epi  wait_weights  issue+backpressure")
tot = [0, 0, 0]
for g, t0, t1, we, ww in out:
    print("  G%d.%d   %7d  %6d   %6d   %6d   %6d" % (g // 10, g % 10, t0, t1 - t0, we, ww, t1 - t0 - we - ww))
    tot[0] += t1 - t0; tot[1] += we; tot[2] += ww
print("tile: first start %d, last end %d; in-half total %d, wait_epi %d, wait_weights %d" % (out[0][1], out[-1][2], tot[0], tot[1], tot[2]))
