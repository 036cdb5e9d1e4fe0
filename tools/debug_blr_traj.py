import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from distbo_b200.data.toy_datasets import one_dim_split_tasks
from distbo_b200.model_embed import set_model
from distbo_b200.train import gp_regressor
from oracle.gp_regressor_oracle import OracleGPRegressor
from oracle import gp_oracle as O

target, sources, supp = one_dim_split_tasks(0, 3, 2, n_train=120, n_test=120)
rs = np.random.RandomState(1)
init = []
for s in sources:
    f, _, _, _ = set_model('toy', s, None)
    th = rs.uniform(-8, 8, size=12)
    init.append(np.column_stack([th, [f(theta=t) for t in th]]))
X = np.vstack(init)[:, :1]; Y = np.vstack(init)[:, 1:]
sizes = [12] * len(sources)
ft, _, _, _ = set_model('toy', target, None)
dX = [np.asarray(s.train_x) for s in sources]; dY = [np.asarray(s.train_y) for s in sources]
common = dict(target_X=np.asarray(target.train_x), target_Y=np.asarray(target.train_y), model_type='regression',
              target_embed_ind=np.arange(len(target.train_x)), float64=True, seed=123)
g, o = gp_regressor('dist', **common), OracleGPRegressor('dist', **common)
fit = dict(data_X=list(dX), data_Y=list(dY), source_embed_ind=[np.arange(len(x)) for x in dX], sizes=list(sizes),
           embed_op='joint', algorithm='BLR', weight_dim=[50, 50, 50], weight_dim_x=[20, 10], weight_dim_y=[20, 10],
           num_of_rff=0, batch_size=500, stratify=False, learning_rate=0.005, max_epochs=25, likhd_var_init=1e-2,
           first_early_stop_epoch=10)
crs = np.random.RandomState(7)
for it in range(6):
    lg = g.maximise_likhd(X, Y, **fit); lo = o.maximise_likhd(X, Y, **fit)
    cand = crs.uniform(-8, 8, size=(20000, 1))
    mg, sg = g.predict_y(cand, compute_embed=True); mo, so = o.predict_y(cand, compute_embed=True)
    ymax = float(Y[sum(sizes[:len(sources)]):].max()) if it else 0.0
    eg, eo = O.acq_ei(mg, sg, ymax, 0.01).ravel(), O.acq_ei(mo, so, ymax, 0.01).ravel()
    top = np.sort(eo)[-3:]
    print(it, "loss rel", abs(lg - lo) / abs(lo), "hist", np.max(np.abs(np.array(g.loss_history) - np.array(o.loss_history)) / np.abs(o.loss_history)),
          "mean rel", np.max(np.abs(mg - mo)) / np.max(np.abs(mo)), "std rel", np.max(np.abs(sg - so)) / np.max(so),
          "argmax", int(np.argmax(eg)), int(np.argmax(eo)), "top3 EI", top, "ei maxdiff", np.max(np.abs(eg - eo)))
    x = cand[np.argmax(eo)]
    yv = ft(theta=float(x[0]))
    X = np.vstack([X, x[None]]); Y = np.vstack([Y, [[yv]]])
    if it == 0:
        fit['data_X'].append(common['target_X']); fit['data_Y'].append(common['target_Y'])
        fit['source_embed_ind'].append(np.arange(len(common['target_X']))); fit['sizes'] = sizes + [1]
    else:
        fit['sizes'] = fit['sizes'][:-1] + [fit['sizes'][-1] + 1]
