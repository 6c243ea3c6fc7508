import sys, numpy as np
sys.path.insert(0, '/root/repo')
import art_b200
from art_b200 import synth
hp = art_b200.HotPath(0)
W,H=3120,2080
xt=synth.xtrans_matrix(); raw=synth.xtrans_frame(W,H,xt,seed=1004)
cam=np.array(synth.XTRANS_RGB_CAM,np.float32)
out=hp.demosaic_xtrans(raw,xt,cam,3,1)
print(float(out[1].mean()))
