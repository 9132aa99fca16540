import sys,os
sys.path.insert(0,'/root/repo')
from dugksfoam_b200 import capi, case as cs
case=cs.poly_cavity_case(120,28)
dv=capi.fvDVM(case, device=0)
print(dv.stats())
dv.close()
