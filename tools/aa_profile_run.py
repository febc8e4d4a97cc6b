import sys, os
sys.path.insert(0, '/root/repo')
from lbmcl_b200.capi import Simulation
with Simulation(dim=256, precision="f32", stride=32, variant=8) as s:
    s.init(); s.run(12, 0); s.sync()
