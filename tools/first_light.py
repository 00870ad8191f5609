import os, time, numpy as np, sys
os.environ['ADAPT_QUIET']='1'
sys.path.insert(0, '/root/repo')
from adapt_b200.parsers.xml_parser import scene_parsing
from adapt_b200._lib import pack_scene
from adapt_b200.renderer.vanilla_renderer import Renderer
from oracle.pt_oracle import OracleScene
root = os.environ.get('SCENES', '/root/repo/scenes')
for scene, name in [('cbox','cbox.xml'), ('csphere','balls-mono.xml')]:
    e,a,o,c = scene_parsing(os.path.join(root, scene), name)
    c['film']['width']=128; c['film']['height']=128
    spp = 16
    rdr = Renderer(e,a,o,c, seed=0)
    t=time.time(); rdr.render_batch(spp); img = rdr.pixels.to_numpy(); dt=time.time()-t
    st = rdr.stats()
    osc = OracleScene(pack_scene(e,a,o,c, seed=0))
    acc, cn = osc.render(spp)
    ref = acc/spp
    rel = np.linalg.norm(img-ref)/np.linalg.norm(ref)
    print(name, 'gpu s', round(dt,3), 'rel L2', rel, 'max abs', np.abs(img-ref).max())
    print('  gpu', {k:st[k] for k in ('paths','rays_closest','rays_shadow','iterations','ms_logic','ms_closest','ms_shadow')})
    print('  cpu', cn)
    # per-pixel mismatch stats
    d = np.abs(img-ref).sum(-1); print('  pixels differing >1e-3:', (d>1e-3).sum(), 'of', d.size)

# all-BxDF coverage scene
e,a,o,c = scene_parsing(os.path.join(root, 'test'), 'allbxdf.xml')
c['film']['width']=128; c['film']['height']=128
spp=16
rdr = Renderer(e,a,o,c, seed=3)
rdr.render_batch(spp); img = rdr.pixels.to_numpy(); st = rdr.stats()
osc = OracleScene(pack_scene(e,a,o,c, seed=3))
acc, cn = osc.render(spp); ref = acc/spp
print('allbxdf rel L2', np.linalg.norm(img-ref)/np.linalg.norm(ref), 'nan', np.isnan(img).sum(), np.isnan(ref).sum(), 'inf', np.isinf(img).sum(), np.isinf(ref).sum())
print('  gpu', {k:st[k] for k in ('paths','rays_closest','rays_shadow','iterations')}); print('  cpu', cn)
d = np.abs(img-ref).sum(-1); print('  pixels differing >1e-3:', (d>1e-3).sum(), 'of', d.size)
# ray batch parity
rng = np.random.default_rng(0)
n=200000
ro = rng.uniform([0.1,0.1,0.1],[5.4,5.4,5.5],(n,3)).astype(np.float32)
rd = rng.normal(size=(n,3)).astype(np.float32); rd/=np.linalg.norm(rd,axis=1,keepdims=True)
g = rdr.intersect_batch(ro, rd); r = osc.intersect_batch(ro, rd)
same = (g['prim']==r['prim'])
print('ray batch: prim equal', same.mean(), 'max |dt| on equal', np.abs(g['t']-r['t'])[same].max(), 'obj equal', (g['obj']==r['obj']).mean())
tm = rng.uniform(0.5,6,(n,)).astype(np.float32)
ga = rdr.intersect_batch(ro, rd, tm, any_hit=True); ra = osc.intersect_batch(ro, rd, tm, any_hit=True)
print('any-hit equal', (ga['prim']==ra['prim']).mean())
