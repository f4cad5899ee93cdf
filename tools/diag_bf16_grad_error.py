"""Gradient error of the native bf16 engine vs fp32 autograd, next to PyTorch bf16 autocast (cuDNN) on the same
small ResNet: shows the native path is within the precision of the dtype (profiles/bf16_grad_error.txt)."""
import sys, torch
import os; R = os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from test_gpu_engine import _resnet
from distkeras_b200.parallel.replica import TorchReplica
from distkeras_b200.parallel.engine import NativeReplica
from distkeras_b200.models.core import compute_loss
B=128
model=_resnet(0)
torch.manual_seed(0)
x=torch.rand(B,16,16,3); y=torch.randint(0,10,(B,))
ref=TorchReplica(model.copy(), {"class_name":"sgd","config":{"lr":0.0}}, "categorical_crossentropy", device="cpu")
ref.train_on_batch(x,y); gref=ref.W.grad.clone()
# bf16 autocast on GPU
m2=model.copy().to("cuda")
W=m2.flat.clone().requires_grad_(True)
with torch.autocast("cuda", dtype=torch.bfloat16):
    out=m2.forward(x.cuda(), flat=W, training=True, logits=True, ctx={})
loss=compute_loss("categorical_crossentropy", out.float(), y.cuda(), True); loss.backward()
gac=W.grad.cpu()
nat=NativeReplica(model, {"class_name":"sgd","config":{"lr":0.0}}, "categorical_crossentropy", B, 0, in_dtype="f32")
nat.train_on_batch(x, y.to(torch.int32)); torch.cuda.synchronize(); gn=nat.G.cpu()
print("%-4s %-22s %9s %9s %9s" % ("li","name","native","autocast","nat-vs-ac"))
for seg in model.segments:
    if not seg.trainable: continue
    sl=slice(seg.offset, seg.offset+seg.size)
    e=lambda a,b: float((a-b).norm()/(b.norm()+1e-12))
    print("%-4d %-22s %9.4f %9.4f %9.4f" % (seg.layer_index, seg.name, e(gn[sl],gref[sl]), e(gac[sl],gref[sl]), e(gn[sl],gac[sl])))
