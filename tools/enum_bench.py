import time, sys, os
sys.path.insert(0,'/root/repo')
from metis_b200 import flatten, native
lib=native.load_library()
best=1e9
for _ in range(5):
    t=time.perf_counter(); sp=flatten.build_plan_space(1,64,512,96,1,6,lib); best=min(best,(time.perf_counter()-t)*1e3)
print(os.environ.get('METIS_ENUM_THREADS'), 'best ms', round(best,2), sp.num_plans)
