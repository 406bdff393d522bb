import sys, os, json, time
os.environ.setdefault('TORCH_CUDA_ARCH_LIST','10.0')
sys.path.insert(0,'/root/reference')
t=time.time()
import torch
from graphs.stylegan_v2_real.networks import Generator
print('import ok', time.time()-t)
out={}
for size in (16,256,1024):
    g=Generator(size,512,8)
    out[size]=[(k,list(v.shape)) for k,v in g.state_dict().items()]
json.dump(out,open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_state_dict_keys.json'),'w'))
print('done')
