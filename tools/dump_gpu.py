"""Renders small configs on the GPU and dumps the mean buffers to gpurun_out/ for offline comparison."""
import os, sys, numpy as np
os.environ['ADAPT_QUIET']='1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapt_b200.parsers.xml_parser import scene_parsing
from adapt_b200.renderer.vanilla_renderer import Renderer
root = 'scenes'
os.makedirs('gpurun_out', exist_ok=True)
out = {}
for tag, scene, name, seed in [('mono','csphere','balls-mono.xml',0), ('all','test','allbxdf.xml',3), ('cbox','cbox','cbox.xml',0)]:
    for spp in (1, 16):
        e,a,o,c = scene_parsing(os.path.join(root, scene), name)
        c['film']['width']=128; c['film']['height']=128
        rdr = Renderer(e,a,o,c, seed=seed)
        rdr.render_batch(spp)
        out[f'{tag}_{spp}'] = rdr.pixels.to_numpy()
        rdr.close()
e,a,o,c = scene_parsing(os.path.join(root, 'test'), 'allbxdf.xml')
c['film']['width']=96; c['film']['height']=96
rdr = Renderer(e,a,o,c, seed=3); rdr.render_batch(16); out['all96_16'] = rdr.pixels.to_numpy(); rdr.close()
np.savez_compressed('gpurun_out/gpu_dump.npz', **out)
print('saved', list(out))
