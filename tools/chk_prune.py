import sys; sys.path.insert(0,'/root/repo')
import os
from pathlib import Path
from ampliconreconstructorom_b200 import capi
ctx=capi.Context(0)
tot=0; worst=2.0
step=1<<28
for lo in range(0,0x7f800001,step):
    n,r=ctx.check_prune_bound(lo,min(lo+step,0x7f800001)); tot+=n; worst=min(worst,r)
import math
print(os.environ.get("SA_LIB_PATH"), "violations",tot,"min exact/bound",repr(worst),"-> needs C >=", -math.log2(worst) if worst<1 else 0)
# per-binade worst ratio in the useful domain
for e in range(0x3a,0x4c):
    n,r=ctx.check_prune_bound(e<<23,(e+1)<<23)
    print(hex(e<<23), n, r)
