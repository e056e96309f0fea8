#!/usr/bin/env python3
"""oracle/mk_ref_inc.py -- TEST INFRASTRUCTURE. Generates forwarding headers so that the reference's MSVC-flavoured
`#include "Utils\\thread_memory.h"` / `<opencv2\\core\\core.hpp>` lines resolve under GCC on Linux WITHOUT touching or copying
any reference source: for every include that contains a backslash (or differs in letter case from the file on disk) reachable
from the given root files, a one-line header literally named like the include text is written into the output directory;
it includes the real reference header by absolute path, or oracle/cvshim/cvshim.hpp for every opencv2 header.

usage: mk_ref_inc.py <outdir> <reference .cpp/.h> ...
"""
import os
import re
import sys

REF = os.environ.get("REF", "/root/reference")
ROOTS = [REF + "/Core/MAGESLAM/Source", REF + "/Dependencies/Arcana/Shared", REF + "/Dependencies/GSL/include",
         REF + "/Dependencies/cereal/include"]
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cvshim", "cvshim.hpp")


def find_ci(rel):
    for root in ROOTS:
        cur = root
        for part in rel.split("/"):
            try:
                hits = [e for e in os.listdir(cur) if e.lower() == part.lower()]
            except OSError:
                hits = []
            if not hits:
                cur = None
                break
            cur = os.path.join(cur, hits[0])
        if cur and os.path.isfile(cur):
            return cur
    return None


def main():
    out = sys.argv[1]
    os.makedirs(out, exist_ok=True)
    seen = set()

    def scan(path):
        if path in seen:
            return
        seen.add(path)
        with open(path, errors="replace") as f:
            txt = f.read()
        for m in re.finditer(r'#\s*include\s*[<"]([^>"]+)[>"]', txt):
            inc = m.group(1)
            fwd = inc.replace("\\", "/")
            is_cv = fwd.lower().startswith("opencv2/")
            local = os.path.join(os.path.dirname(path), fwd)
            tgt = None if is_cv else (local if os.path.isfile(local) else find_ci(fwd))
            if is_cv or (tgt and "\\" in inc):
                with open(os.path.join(out, inc), "w") as f:
                    f.write('#include "%s"\n' % (SHIM if is_cv else tgt))
            elif tgt and not os.path.isfile(local) and not any(os.path.isfile(os.path.join(r, fwd)) for r in ROOTS):
                name = os.path.join(out, fwd)                      # letter-case mismatch only
                os.makedirs(os.path.dirname(name), exist_ok=True)
                with open(name, "w") as f:
                    f.write('#include "%s"\n' % tgt)
            if tgt:
                scan(tgt)

    for r in sys.argv[2:]:
        scan(r)
    print("mk_ref_inc: %d reference files scanned, forwarding headers in %s" % (len(seen), out))


if __name__ == "__main__":
    main()
