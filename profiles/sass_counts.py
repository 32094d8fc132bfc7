#!/usr/bin/env python
"""Per-object SASS opcode counts of the built library (evidence that the tcgen05 / TMA / mbarrier paths are what was compiled):
    python profiles/sass_counts.py > profiles/r02_sass_opcode_counts.json
Reads manifold_gp_b200/csrc/_obj/*.o (built by csrc/build.py for sm_100a) with `cuobjdump -sass`."""
import json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "manifold_gp_b200", "csrc", "_obj")
OPS = {"UTCHMMA": "tcgen05.mma (TF32 tensor-core MMA)", "UTMALDG": "cp.async.bulk.tensor (TMA tile load)", "LDTM": "tcgen05.ld (TMEM -> registers)",
       "UTCBAR": "tcgen05.commit (MMA completion -> mbarrier)", "UTCATOMSWS": "tcgen05.alloc / dealloc",
       "UBLKCP": "cp.async.bulk (TMA bulk copy)", "LDGSTS": "cp.async (16-byte global -> shared)",
       "SYNCS": "mbarrier arrive / try_wait / expect_tx", "SHFL": "warp shuffle", "REDUX": "warp reduce",
       "ATOM": "global atomics", "LDS": "shared loads", "STS": "shared stores", "HMMA": "legacy mma.sync (none expected)",
       "FFMA": "fp32 FMA", "DFMA": "fp64 FMA"}
out = {}
for f in sorted(os.listdir(OBJ)):
    if not f.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, f)], capture_output=True, text=True).stdout
    row = {}
    for op in OPS:
        c = len(re.findall(r"\b" + op + r"[\w.]*", sass))
        if c:
            row[op] = c
    row["kernels"] = len(re.findall(r"Function : ", sass))
    out[f] = row
print(json.dumps({"legend": OPS, "objects": out, "arch": "sm_100a", "how": "cuobjdump -sass per object, regex count of opcode mnemonics"}, indent=1))
