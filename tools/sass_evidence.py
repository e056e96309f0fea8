#!/usr/bin/env python3
"""Counts the Blackwell-only SASS mnemonics per kernel of libmage_b200.so (tcgen05 MMA / TMEM load / commit, TMA tensor load, mbarrier ops):
the evidence B200_PROFILING.md asks for that a kernel really uses the tensor cores / TMA. usage: sass_evidence.py [lib] > profiles/rNN_sass_evidence.md"""
import collections, os, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mageslam_b200", "libmage_b200.so")
PAT = re.compile(r"UTCIMMA|UTCHMMA|UTCQMMA|LDTM|STTM|UTCBAR|UTCATOMSWS|UTMALDG|UBLKCP|SYNCS|FENCE\.VIEW\.ASYNC|PREEXIT|ACQBULK")


def collect(path):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m and cur and PAT.match(m.group(1)):
            op = m.group(1)
            cnt[cur][op if op.startswith("SYNCS") else op.split(".")[0]] += 1
    return cnt


if __name__ == "__main__":
    cnt = collect(lib)
    print("# Blackwell-specific SASS per kernel (`cuobjdump -sass %s`)\n" % os.path.basename(lib))
    print("`UTCIMMA` = tcgen05.mma kind::i8, `LDTM` = tcgen05.ld, `UTCBAR` = tcgen05.commit, `UTCATOMSWS` = tcgen05.alloc / dealloc, "
          "`UTMALDG` = cp.async.bulk.tensor (TMA), `SYNCS.*` = mbarrier operations, `FENCE.VIEW.ASYNC` = fence.proxy.async, "
          "`PREEXIT` / `ACQBULK` = griddepcontrol.launch_dependents / .wait (programmatic dependent launch).\n")
    print("| kernel | mnemonics (static count) |\n|---|---|")
    for k, v in sorted(cnt.items()):
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
        print("| `%s` | %s |" % (name, ", ".join("%s ×%d" % kv for kv in sorted(v.items()))))
