"""Static SASS mnemonic counts per kernel of the built library (no GPU needed):
    python tools/sass_mnemonics.py [mtgs_b200/libb200splat.so] > profiles/rNN_sass_mnemonics.txt
Only the instantiations MTGS reaches are listed for the templated blend kernels (CDIM 4 with 3/4 channels, CDIM 8 with
6/7 channels); everything else is listed once per kernel."""
import collections
import re
import subprocess
import sys

WANT = ["FFMA2", "FMUL2", "FADD2", "UBLKCP", "SYNCS", "REDG", "RED.", "ATOMS", "ATOMG", "MATCH", "REDUX", "VOTE", "SHFL",
        "MUFU.EX2", "MUFU.RCP", "MUFU.LG2", "BAR.SYNC", "UTMACMDFLUSH", "CCTL", "LDS.128", "STS.128", "LDG.E.128", "STG.E.128",
        "NANOSLEEP", "MEMBAR", "ERRBAR"]


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True,
                           text=True).stdout.splitlines()
    counts, order, cur, k = {}, [], None, -1
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            k += 1
            cur = re.sub(r"\(.*", "", names[k])
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for w in WANT:
                if op.startswith(w):
                    counts[cur][w.rstrip(".")] += 1
                    break
    print("# static SASS mnemonic counts per kernel of", path, "(cuobjdump -sass, sm_100a)")
    print("# FFMA2/FMUL2/FADD2 = packed fp32x2 arithmetic; UBLKCP = TMA bulk copy (cp.async.bulk, global<->shared);")
    print("# SYNCS = mbarrier ops; REDG = reductions to global memory (incl. red.global.add.v4.f32); REDUX = redux.sync;")
    print("# MATCH = match.any; ATOMS / ATOMG = shared / global atomics; NANOSLEEP = the exchange's flag spin loops")
    for name in sorted(order):
        if re.search(r"k_blend_(fwd|bwd)<", name) and not re.search(r"<4, [34], |<8, [67], ", name):
            continue
        c = counts[name]
        body = "  ".join(f"{w}:{c[w]}" for w in sorted(c) if w != "_total")
        print(f"{name:60s} n={c['_total']:5d}  {body}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "mtgs_b200/libb200splat.so")
