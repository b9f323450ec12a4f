import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from slam_pose_estimation_b200 import UkfBatch, synthetic as syn
B=1<<20
mu,sg=syn.pose_initial(B,perturb=True)
dev=torch.device('cuda:0')
for kern in ('fast','thread'):
    os.environ['UKFB_KERNEL']=kern
    f=UkfBatch(0,B); f.initialize(mu,sg)
    d_dt=torch.full((1,),syn.DT,dtype=torch.float64,device=dev)
    for kind in (8,3):
        z,R=syn.pose_measurement(kind,B,1)
        if kind==3:
            q=mu[:,3:7]; n=np.linalg.norm(q[:,:3],axis=1); ang=2*np.arctan2(n,q[:,3]); z=q[:,:3]/np.maximum(n,1e-300)[:,None]*ang[:,None]+z*0.1
        d_z=torch.from_numpy(np.ascontiguousarray(z)).to(dev); d_R=torch.from_numpy(R).to(dev)
        for _ in range(3): f.step_dev(d_dt,False,kind,d_z,d_R,False)
        f.synchronize(); f.event_record(0)
        for _ in range(10): f.step_dev(d_dt,False,kind,d_z,d_R,False)
        f.event_record(1); f.synchronize()
        ms=f.event_elapsed_ms(0,1)/10
        print(kern,'kind',kind,round(ms,3),'ms',round(B/ms/1e3,1),'M/s flagged',f.status_summary())
    f.close()
