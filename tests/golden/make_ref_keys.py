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
# discriminator layout
from graphs.stylegan_v2_real.networks import Discriminator
outd = {}
for size, cm in ((32, 2), (64, 1), (256, 2), (1024, 2)):
    d = Discriminator(size, channel_multiplier=cm)
    outd[f"{size}x{cm}"] = [(k, list(v.shape)) for k, v in d.state_dict().items()]
json.dump(outd, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_discriminator_keys.json'), 'w'))
print('discriminator keys done')
