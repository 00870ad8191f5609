"""Offline: compare gpurun_out/gpu_dump.npz against the CPU oracle (same seeds / sizes as tools/dump_gpu.py)."""
import numpy as np, os, sys
os.environ['ADAPT_QUIET']='1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapt_b200.parsers.xml_parser import scene_parsing
from adapt_b200._lib import pack_scene
from oracle.pt_oracle import OracleScene
g = np.load('gpurun_out/gpu_dump.npz')
show = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for tag, scene, name, seed in [('mono','csphere','balls-mono.xml',0), ('all','test','allbxdf.xml',3), ('cbox','cbox','cbox.xml',0)]:
    e,a,o,c = scene_parsing(os.path.join('scenes', scene), name)
    c['film']['width']=128; c['film']['height']=128
    osc = OracleScene(pack_scene(e,a,o,c, seed=seed))
    for spp in (1,16):
        acc,_ = osc.render(spp); ref = acc/spp
        img = g[f'{tag}_{spp}']
        d = np.abs(img-ref).sum(-1)
        bad = np.argwhere(d > 1e-3*np.maximum(1, np.abs(ref).sum(-1)))
        good = d <= 1e-3*np.maximum(1, np.abs(ref).sum(-1))
        relg = np.linalg.norm((img-ref)[good])/np.linalg.norm(ref[good])
        print(tag, 'spp', spp, 'rel L2 %.3e' % (np.linalg.norm(img-ref)/np.linalg.norm(ref)), 'bad pixels', len(bad), 'rel L2 over matching %.2e' % relg,
              'nan', int(np.isnan(img).sum()), int(np.isnan(ref).sum()))
        if spp == 1:
            for (i,j) in bad[:show]:
                print('   ', i,j, img[i,j], ref[i,j])
