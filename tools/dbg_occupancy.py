import sys, os, subprocess, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/dispersal-simulation_b200/python')
import bench, dsim_b200 as D, ctypes as C
F="index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
r=subprocess.run(["nvidia-smi",f"--query-gpu={F}","--format=csv,noheader,nounits","-i","0"],capture_output=True,text=True); print("SMI:",r.returncode,r.stdout,r.stderr)
lib=D.load(); synth=D.load_synth()
g,res,mx,fam=bench.build_workload(bench.CONFIGS["c2"],1,synth)
sim=D.Sim(lib,g,seed=7); sim.set_patches(res,mx); sim.set_families(fam)
for k in range(4):
    st=sim.step(50)
    f=sim.get_families(history=False,adaptation=False)
    cnt=np.unique(f.location,return_counts=True)[1]
    h=np.bincount(np.minimum(cnt,40))
    print("t",st.t,"live",st.live_families,"occ",len(cnt),"max",cnt.max(),"hist(1..12)",h[1:13],">8:",(cnt>8).sum(),">32:",(cnt>32).sum(), "fam in >8:", cnt[cnt>8].sum())
