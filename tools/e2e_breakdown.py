import os, sys, time
import numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import common
from pupperv3_mjx_b200 import abi, runtime, prng, domain_randomization as dr
n=4096
env=common.make_env(); env.set_episode_params(1000,1)
def fresh():
    rt=runtime.EnvRuntime(env.model_desc, env.env_cfg, n, episode=True)
    sv,_=dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), n)); rt.set_dr(sv)
    rt.reset(torch.from_numpy(np.ascontiguousarray(prng.split(prng.PRNGKey(0), n)).view(np.int32)).cuda())
    return rt
w=env.env_cfg.observation_history*abi.OBS_DIM
acts=[torch.from_numpy(common.actions(n,t)) for t in range(8)]
h_act=[a.pin_memory() for a in acts]; d_act=[a.cuda() for a in acts]
h_out=torch.empty(n*(w+2)).pin_memory()
# device path kernel time (events), L2 hot
rt=fresh()
for t in range(20): rt.step(d_act[t%8])
ev=[(torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)) for _ in range(200)]
for t in range(200):
    ev[t][0].record(); rt.step(d_act[t%8]); ev[t][1].record(); torch.cuda.current_stream().synchronize()
print("device path kernel (events, sync each step): %.1f us"%(1e3*np.median([a.elapsed_time(b) for a,b in ev])))
rt=fresh()
for t in range(20): rt.step_host(h_act[t%8],h_out).synchronize()
tl=[];ts=[];tk=[]
for t in range(200):
    ev[t][0].record()
    t0=time.perf_counter(); s=rt.step_host(h_act[t%8],h_out); t1=time.perf_counter(); ev[t][1].record(); s.synchronize(); t2=time.perf_counter()
    tl.append(t1-t0); ts.append(t2-t1)
torch.cuda.synchronize()
print("host path: kernel (events) %.1f us, python call %.1f us, wait %.1f us, total %.1f us"%(1e3*np.median([a.elapsed_time(b) for a,b in ev]), 1e6*np.median(tl), 1e6*np.median(ts), 1e6*np.median(np.array(tl)+np.array(ts))))
