import ctypes, os, sys
from pathlib import Path
sys.path.insert(0, "."); sys.path.insert(0, "benchmarks")
os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")
import torch
import profile_step as ps
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics import Accuracy
from fl4health_b200.clients.basic_client import BasicClient
rt = ctypes.CDLL("libcudart.so.12")
def status(tag):
    st = ctypes.c_int(-1)
    rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    if st.value != 0: print("   ", tag, "status", st.value, flush=True)
class Dbg(ps.Client):
    def train_step(self, input, target):
        opt = self.optimizers["global"]; opt.zero_grad(); status("zero_grad")
        try:
            preds, feats = self.predict(input); status("forward")
        except Exception:
            import traceback; traceback.print_exc(limit=12); raise
        losses = self.compute_training_loss(preds, feats, target); status("loss")
        losses.backward["backward"].backward(); status("backward")
        opt.step(); status("opt.step")
        return losses, preds
device = torch.device("cuda:0")
engine = EngineOptions(cuda_graphs=True, amp_dtype=None, channels_last=True, master_weights=False,
                       table_grads=os.environ.get("FL4H_TABLE_GRADS", "1") != "0")
c = Dbg(Path("."), [Accuracy()], device, client_name="dbg", engine_options=engine)
c.setup_client({"current_server_round": 1, "local_steps": 8, "batch_size": ps.BS})
c.model.train()
x, y = c._prepare_batch(*next(iter(c.train_loader)))
print("input", x.shape, x.stride(), x.dtype, "w", c.model.conv1.weight.stride(), c.model.conv1.weight.is_contiguous(memory_format=torch.channels_last))
for i in range(6):
    c._run_train_unit(x, y)
torch.cuda.synchronize()
print("replays", getattr(c._train_runner, "replays", None))
