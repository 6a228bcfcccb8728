#!/usr/bin/env python
"""Fingerprint of the device code that is actually shipped: md5 of the SASS text (addresses and encodings stripped) of every
object of the in-tree build.  `--write` records it in profiles/sass_fingerprint.json next to the GPU run that verified it;
without arguments the current build is compared with the record -- an edit that leaves every fingerprint unchanged did not
change what runs on the device, so the last GPU verification still stands.

    python __graft_entry__.py && python tools/sass_fingerprint.py [--write "note about the verifying run"]"""
import hashlib, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "build", "obj")
REC = os.path.join(ROOT, "profiles", "sass_fingerprint.json")


def fingerprint(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    h = hashlib.md5()
    for ln in out.splitlines():
        if re.match(r"^\s+/\*[0-9a-f]{4}\*/", ln):
            h.update(re.sub(r"/\*[0-9a-f]+\*/", "", ln).encode() + b"\n")
    return h.hexdigest()


cur = {f: fingerprint(os.path.join(OBJ, f)) for f in sorted(os.listdir(OBJ)) if f.endswith(".o")}
if len(sys.argv) > 1 and sys.argv[1] == "--write":
    json.dump({"verified_by": " ".join(sys.argv[2:]), "objects": cur}, open(REC, "w"), indent=1)
    print("recorded", REC)
    sys.exit(0)
rec = json.load(open(REC))
same = True
for f, h in cur.items():
    ok = rec["objects"].get(f) == h
    same &= ok
    print(f"{f:24s} {h}  {'== verified' if ok else '!= verified (' + str(rec['objects'].get(f)) + ')'}")
print("device code identical to the build verified by:" if same else "DEVICE CODE CHANGED since the build verified by:", rec["verified_by"])
sys.exit(0 if same else 1)
