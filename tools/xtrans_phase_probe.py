import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import art_b200
from art_b200 import synth
hp = art_b200.HotPath(0)
W,H=6240,4160
xt=synth.xtrans_matrix(); raw=synth.xtrans_frame(W,H,xt,seed=1004)
cam=np.array(synth.XTRANS_RGB_CAM,np.float32)
for stop in (1,2,3,4,5,0):
    os.environ["ART_XT_STOP"]=str(stop)
    hp.profile_enable(True)
    for i in range(3): out=hp.demosaic_xtrans(raw,xt,cam,3,1)
    print("stop",stop,{k:round(v[0]/v[1],3) for k,v in hp.profile_collect().items() if k=="k_xtrans"})
