"""Host cost of cudaGraphLaunch for the REAL captured train / val steps in the patterns of an FL round (see
graph_launch_probe.py for the synthetic version): steady state, first after idle, first after the other graph ran,
and with an early cudaGraphUpload."""
import statistics, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import torch
from cuda.bindings import runtime as rt
import profile_step as ps
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics import Accuracy

device = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
engine = EngineOptions(cuda_graphs=True, amp_dtype=torch.bfloat16, channels_last=True, master_weights=True)
client = ps.Client(Path("."), [Accuracy()], device, client_name="prof", engine_options=engine)
client.setup_client({"current_server_round": 1, "local_steps": 8, "batch_size": ps.BS})
x, y = next(iter(client.train_loader))
x, y = client._prepare_batch(x, y)
client.model.train()
for _ in range(8):
    client._run_train_unit(x, y)
client.model.eval()
for _ in range(8):
    client._run_val_unit(x, y, client.val_loss_meter, client.val_metric_manager)
torch.cuda.synchronize()
A = next(iter(next(iter(client._train_runners.values()))._graphs.values())).graph
B = next(iter(next(iter(client._val_runners.values()))._graphs.values())).graph
stream = torch.cuda.current_stream()


def t(fn) -> float:
    t0 = time.perf_counter()
    fn()
    return (time.perf_counter() - t0) * 1e6


def upload(g) -> None:
    (err,) = rt.cudaGraphUpload(g.raw_cuda_graph_exec(), stream.cuda_stream)
    assert err == rt.cudaError_t.cudaSuccess, err


keys = ("A busy", "A first after idle (A was last)", "A after B, GPU busy", "A first after B, idle", "B first after A, idle",
        "upload(A) host cost", "A first after B, idle, uploaded early", "B first after A, idle, uploaded early",
        "A first after B, idle, upload right before launch")
res = {k: [] for k in keys}
for _ in range(20):
    A.replay()
    res["A busy"].append(t(A.replay))
    torch.cuda.synchronize()
    res["A first after idle (A was last)"].append(t(A.replay))
    for _ in range(4):
        B.replay()
    res["A after B, GPU busy"].append(t(A.replay))
    for _ in range(4):
        B.replay()
    torch.cuda.synchronize()
    res["A first after B, idle"].append(t(A.replay))
    for _ in range(7):
        A.replay()
    torch.cuda.synchronize()
    res["B first after A, idle"].append(t(B.replay))
    for _ in range(3):
        B.replay()
    res["upload(A) host cost"].append(t(lambda: upload(A)))
    torch.cuda.synchronize()
    res["A first after B, idle, uploaded early"].append(t(A.replay))
    for _ in range(7):
        A.replay()
    upload(B)
    torch.cuda.synchronize()
    res["B first after A, idle, uploaded early"].append(t(B.replay))
    for _ in range(3):
        B.replay()
    torch.cuda.synchronize()
    upload(A)
    res["A first after B, idle, upload right before launch"].append(t(A.replay))
    torch.cuda.synchronize()
for k, v in res.items():
    print(f"{k:52s} median {statistics.median(v):8.1f} us   min {min(v):8.1f}  max {max(v):8.1f}")
