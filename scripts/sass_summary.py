"""SASS mnemonic counts per kernel of probnmn_clevr_b200/libpnmn.so (cuobjdump -sass): which kernels carry tcgen05 MMAs
(UTCHMMA / UTCQMMA), TMEM loads (LDTM), bulk async copies (UBLKCP), tensor-map TMA (UTMALDG), mbarrier ops (SYNCS), vector
reductions (REDG ... F32x4).  Usage: python scripts/sass_summary.py > profiles/r2_sass.md"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "probnmn_clevr_b200", "libpnmn.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
keys = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "REDG", "F32x4", "MUFU", "BAR.SYNC"]
cur, rows, k = None, collections.OrderedDict(), 0
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names[k][:70]; k += 1
        rows[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m:
        op = m.group(1)
        rows[cur]["n"] += 1
        for key in keys:
            if key in op:
                rows[cur][key] += 1
print("# SASS summary of libpnmn.so (sm_100a), `python scripts/sass_summary.py`\n")
print("| kernel | instr | " + " | ".join(keys) + " |")
print("|---|---:|" + "---:|" * len(keys))
for name, c in rows.items():
    print(f"| `{name}` | {c['n']} | " + " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")
