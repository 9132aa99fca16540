import sys,os
sys.path.insert(0,'/root/repo')
import bench
from dugksfoam_b200 import capi
case=bench.build_case(*bench.WORKLOADS['cavity3d_64_gh28'])
dv=capi.fvDVM(case, device=0)
print(dv.stats())
dv.close()
