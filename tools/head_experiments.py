import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
code = r'''
import os, sys, torch, json
sys.path.insert(0, %r)
import samplenet_b200 as sb
from tools.kernel_times import graph_time
torch.manual_seed(0)
net = sb.SampleNet(64, 128, group_size=8, input_shape="bnc", output_shape="bnc").cuda().train()
x = torch.rand(32, 1024, 3, device="cuda") - 0.5
conv, fc = net._layer_specs()
with torch.no_grad():
    net(x)
    t = graph_time(lambda: sb.ops.generator_forward(x, "bnc", conv, fc, True, 64, _profile_flags=4))
print(json.dumps({"dbg": os.environ.get("SNB200_HEAD_DEBUG", "0"), "head_only_us": t}))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for dbg in ("0", "1", "2", "4", "8", "5", "9"):
    env = dict(os.environ, SNB200_HEAD_DEBUG=dbg)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-300:], flush=True)
