"""Host cost of cudaGraphLaunch in the patterns an FL round produces: back-to-back replays of one graph, the first
replay after an idle GPU, the first replay after ANOTHER graph ran, and the same with cudaGraphUpload issued early."""
import statistics, time
import torch
from cuda.bindings import runtime as rt

dev = torch.device("cuda:0")
bufs = [torch.zeros(4096, device=dev) for _ in range(160)]
big = torch.zeros(64 << 20, device=dev)


def make(n: int, heavy: bool = False) -> torch.cuda.CUDAGraph:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for b in bufs[:n]:
            b.add_(1.0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for b in bufs[:n]:
                b.add_(1.0)
            if heavy:
                big.add_(1.0)  # ~80 us of GPU work so the launches queue up behind it
    return g


A, B = make(150, heavy=True), make(50, heavy=True)
stream = torch.cuda.current_stream()


def t(fn) -> float:
    t0 = time.perf_counter()
    fn()
    return (time.perf_counter() - t0) * 1e6


def upload(g: torch.cuda.CUDAGraph) -> None:
    (err,) = rt.cudaGraphUpload(g.raw_cuda_graph_exec(), stream.cuda_stream)
    assert err == rt.cudaError_t.cudaSuccess, err


for _ in range(5):
    A.replay(); B.replay()
torch.cuda.synchronize()
res = {k: [] for k in ("A busy", "A first after idle (A was last)", "A first after B, idle", "A first after B, idle, uploaded early",
                       "B first after A, idle", "B first after A, idle, uploaded early", "upload(A) host cost", "A after B, GPU busy")}
for _ in range(30):
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
    torch.cuda.synchronize()
for k, v in res.items():
    print(f"{k:45s} median {statistics.median(v):8.1f} us   min {min(v):8.1f}  max {max(v):8.1f}")
