#!/usr/bin/env python
"""SHA-256 of the SASS instruction stream of every kernel in libbtcdet_b200.so (addresses, encodings and line info
stripped), to show that source edits made after the last GPU run of a round (comments, host code, new template
parameters with a default) did not change the machine code that passed `pytest -m gpu`.

  python tools/sass_fingerprint.py --write profiles/verified_sass.json     # right after a green GPU suite
  python tools/sass_fingerprint.py --check profiles/verified_sass.json     # any time later (CPU only)
Template arguments that were added later with a default value are matched by prefix (`--check` strips a trailing
default when the exact name is missing)."""
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "btcdet_b200", "libbtcdet_b200.so")


def fingerprints():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res, name, lines = {}, None, []

    def flush():
        if name is not None:
            res[name] = {"sha256": hashlib.sha256("\n".join(lines).encode()).hexdigest(), "instructions": len(lines)}
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            flush()
            name, lines = m.group(1), []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m and name is not None:
            lines.append(re.sub(r"\s+", " ", m.group(1)).strip())
    flush()
    return res


def demangled(names):
    out = subprocess.run(["cu++filt"] + list(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}


def kernel_key(pretty):
    """Demangled name without the parameter list: `void btc::k<(int)64, (bool)0>(float const*, ...)` -> `btc::k<64, 0>`."""
    i = pretty.find(">(")
    head = pretty[:i + 1] if i >= 0 else pretty.split("(")[0]
    head = re.sub(r"\((?:int|bool|unsigned int)\)", "", head)
    return head.replace("void ", "").strip()


def current():
    fp = fingerprints()
    pretty = demangled(list(fp))
    return {kernel_key(pretty[k]): v for k, v in fp.items()}


def main():
    mode, path = sys.argv[1], sys.argv[2]
    cur = current()
    if mode == "--write":
        json.dump({"lib": "btcdet_b200/libbtcdet_b200.so", "kernels": cur}, open(path, "w"), indent=1, sort_keys=True)
        print("wrote %d kernel fingerprints to %s" % (len(cur), path))
        return 0
    ref = json.load(open(path))["kernels"]
    bad = 0
    for name, want in sorted(ref.items()):
        got = cur.get(name)
        if got is None:   # a template parameter with a default was appended later: match "<...>" by prefix
            stem = name[:-1] if name.endswith(">") else name
            cands = [k for k in cur if k.startswith(stem + ",")]
            got = next((cur[k] for k in cands if cur[k]["sha256"] == want["sha256"]), None)
        if got is None or got["sha256"] != want["sha256"]:
            bad += 1
            print("CHANGED or missing: %s" % name)
    print("%d of %d verified kernels unchanged" % (len(ref) - bad, len(ref)))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
