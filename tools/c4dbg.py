"""Image / final_T deviation from the oracle on the deep-stack configurations (C4 dense, C5 1080p)."""
import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from guassianhand_b200 import scenes
import util
bg0=np.zeros(3,np.float32)
def report(name, sc, cam):
    gout,_,info = util.run_gpu(sc,[cam],bg0)
    f,_ = util.run_oracle(sc,cam,bg0)
    d = np.abs(f['out_color']-gout[0]['out_color'])
    i = np.unravel_index(d.argmax(), d.shape)
    amb = f['ambig'] != 0
    dn = np.where(amb[None], 0, d); print('   non-ambig maxdiff', dn.max(), 'ambig maxdiff', np.where(amb[None], d, 0).max(), 'finalT nonambig', np.where(amb,0,np.abs(f['final_T']-gout[0]['final_T'])).max())
    print(name,'R',info['R'],'maxdiff',d.max(),'n_contrib there', f['n_contrib'][i[1],i[2]], 'max n_contrib', f['n_contrib'].max(),
          'finalT diff', np.abs(f['final_T']-gout[0]['final_T']).max(), 'n_contrib equal', np.array_equal(f['n_contrib'][~amb],gout[0]['n_contrib'][~amb]), 'ambig', amb.mean())
report('c4dense', scenes.two_hand_scene(1000000, seed=0, sh_degree=3), scenes.fibonacci_cameras(2, 1024, 1024, seed=0)[0])
sc = scenes.two_hand_scene(60000, seed=0)
for cam in scenes.fibonacci_cameras(4, 1080, 1920, seed=2)[:2]:
    report('c5', sc, cam)
report('c2', sc, scenes.fibonacci_cameras(2, 512, 334, seed=0)[0])
