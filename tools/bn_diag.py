import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, argparse
from oracle.gen_golden import seeded_inputs, seeded_out_labels
from oracle.seeded import seeded_state_dict, strided_sample
from oracle import loss as ol
from inferbiomechanics_b200.loss.RegressionLossEvaluator import RegressionLossEvaluator
from inferbiomechanics_b200.models.FeedForwardRegressionBaseline import FeedForwardBaseline
Q = (ol.COP, ol.FORCE, ol.TORQUE, ol.WRENCH)
ALL = argparse.Namespace(predict_grf_components=list(range(6)), predict_cop_components=list(range(6)),
                         predict_moment_components=list(range(6)), predict_wrench_components=list(range(12)))
g=np.load('/root/repo/tests/golden/ff_bn_train.npz')
for name in ["sigmoid_b16","relu_b300"]:
    act=name.split("_")[0]
    D, T, s, B, seed, iseed, lseed = (int(v) for v in g[f"{name}/meta"])
    hidden=[int(v) for v in g[f"{name}/hidden"]]
    m = FeedForwardBaseline(D, 2, T, "all_frames", act, s, 10, hidden_dims=hidden, batchnorm=True)
    m.load_state_dict(seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed))
    m=m.cuda(); m.train()
    inputs = seeded_inputs(B, T//s, D, s * 3, iseed)
    _, labels = seeded_out_labels(B, T//s, lseed)
    out=m(inputs)
    for k in Q:
        r=torch.as_tensor(g[f"{name}/out/{k}"]).double(); o=out[k].detach().double().cpu()
        print(name,k,"out err/scale",((o-r).abs().max()/r.abs().max()).item())
    ev = RegressionLossEvaluator(dataset=None, split="train", device="cuda")
    loss = ev(inputs, out, {k: v.clone() for k, v in labels.items()}, [], [], ALL)
    print("loss", loss.item(), float(g[f"{name}/loss"]))
    loss.backward()
    for n,p in m.named_parameters():
        r=torch.as_tensor(g[f"{name}/grad_sample/{n}"]).double(); o=strided_sample(p.grad).double().cpu()
        print(name,n,"grad err/scale",((o-r).abs().max()/r.abs().max()).item())
    for k,v in m.state_dict().items():
        if "running_" in k:
            r=torch.as_tensor(g[f"{name}/buffer/{k}"]).double(); o=v.double().cpu()
            print(name,k,"buf err/scale",((o-r).abs().max()/r.abs().max()).item())
