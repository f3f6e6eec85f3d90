"""SASS evidence for the Blackwell-specific instructions of the hot kernels:  python profiles/sass_evidence.py > profiles/r02_sass_evidence.txt
Counts per kernel (cuobjdump -sass of csrc/libslb.so, sm_100a): DMMA (FP64 tensor-core m8n8k4 -- FP64 has no tcgen05 type),
UBLKCP (TMA bulk copies, cp.async.bulk), SYNCS (mbarrier), LDGSTS (cp.async), DFMA, LDS / STS, and the register / spill
figures ptxas reported."""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "slam-localization_b200", "csrc", "libslb.so")
OPS = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "WARPSYNC"]
HOT = ["usckf_step_kernel", "predict12_kernel", "ukf_kernel", "msckf_update_kernel", "msckf_ekf_update_kernel", "datamodel_kernel",
       "ekf_update_kernel", "ekf_predict_kernel", "safe_fusion_kernel", "dr_update_pose_kernel", "transform_compose_kernel",
       "check_sigma_points_kernel"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        kern[cur][m.group(1).split(".")[0]] += 1
        kern[cur]["_total"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass %s (sm_100a); static instruction counts per kernel" % os.path.relpath(LIB, ROOT))
print("%-96s %6s " % ("kernel", "instr") + " ".join("%7s" % o for o in OPS))
for (name, c), dn in zip(kern.items(), demangle):
    if not any(h in dn for h in HOT):
        continue
    short = re.sub(r"\(slb::FilterArgs\)|slbd::|void ", "", dn)[:96]
    print("%-96s %6d " % (short, c["_total"]) + " ".join("%7d" % c[o] for o in OPS))
