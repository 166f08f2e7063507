timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "render" 2>&1 | tail -4
python - <<'PY' 2>&1 | tail -4
import torch, sys
sys.path.insert(0,'.')
import bench as B, gridworld_b200 as gw
dev=torch.device('cuda',0)
for n in (16384, 65536):
    env=gw.make_vec(n, device=dev, render=True, max_steps=250)
    T=B.synth_targets(n,0.2,'compact',1).to(dev); env.set_tasks(T, T)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    env.render(); e0.record()
    for _ in range(5): env.render()
    e1.record(); torch.cuda.synchronize()
    print('render', n, 'envs:', e0.elapsed_time(e1)/5, 'ms per frame batch')
    del env
PY
