import sys, torch
sys.path.insert(0, '.')
from oracle import models as om
from tf_repos_b200 import synth
from tf_repos_b200.din import DIN
B, N, K, Fp, P = 64, 5000, 8, 11, 9
opt = sys.argv[1] if len(sys.argv) > 1 else "Adam"
kw = dict(deep_layers="16,8", dropout="1.0,1.0", attention_layers="256", attention_pooling=True, l2_reg=1e-4,
          learning_rate=5e-4 if opt == "Adam" else 0.01, optimizer=opt)
ref = om.DIN(Fp, N, K, update_mode="exact", seed=4, **kw)
g = torch.Generator().manual_seed(1)
ref.params["embeddings"].copy_(torch.randn(N, K, generator=g) * 0.1)
gpu = DIN(Fp, N, K, B, P, max_a_int=8, update_mode="exact", device="cuda:0", **kw)
gpu.load_variables(ref.params)
for step in range(3):
    batch, labels = synth.din_batch(B, N, Fp, P, 8, seed=50 + step)
    lb = {k: (v.long() if v.dtype == torch.int32 else v) for k, v in batch.items()}
    loss, out, tg, dg = ref.gradients(lb, labels)
    cb = {k: v.cuda() for k, v in batch.items()}
    # run gpu fwd/bwd without applying: replicate train_step pieces
    gpu._stage_ids(cb); gpu.opt.tick()
    from tf_repos_b200 import ops
    y_d = gpu._forward(cb, train=True)
    ops.logit_loss(None, y_d, None, None, labels.cuda(), B, y=gpu.y, pred=gpu.pred, loss_ce=gpu.loss_ce, dy=gpu.dy)
    gpu._backward(cb)
    print("step", step, "logit maxdiff", (gpu.y.cpu() - out["y"]).abs().max().item())
    for n, gr in dg.items():
        gg = gpu.dense.grads[n].cpu()
        print("   %-50s |g| %.3e  maxdiff %.3e" % (n, gr.abs().max().item(), (gg - gr).abs().max().item()))
    gpu.updater.dedup(gpu.ids_all, gpu.g_all, None)
    summed, uniq = tg["embeddings"]
    U = gpu.updater.uw.n_uniq.item()
    gu = gpu.updater.g_uniq[:U * K].view(U, K).cpu()
    print("   table grads: U", U, uniq.numel(), "maxdiff %.3e of |g| %.3e" % ((gu - summed).abs().max().item(), summed.abs().max().item()))
    gpu.updater.apply(gpu.V, None, exact=True, l2_reg=1e-4); gpu.dense.apply()
    ref.apply_gradients(tg, dg)
    for n, p in ref.params.items():
        v = gpu.variables()[n].cpu()
        print("   var %-46s scale %.3e maxdiff %.3e" % (n, p.abs().max().item(), (v - p).abs().max().item()))
