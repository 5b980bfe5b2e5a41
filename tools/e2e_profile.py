"""Where the wall time of cbas.encode_file goes on the bench clip (35 x 512 frames of 256x256, .npy in /dev/shm):
cProfile of the calling thread + a few direct timers.  usage: e2e_profile.py [chunks]"""
import cProfile, io, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cbas_b200 import cbas, gui_state
from cbas_b200.encoder import DinoEncoder

n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 35
enc = DinoEncoder("synthetic:vitb16", "cuda", preprocess="processor", image_size=224, max_frames=512)
td = tempfile.mkdtemp(prefix="cbas_prof_", dir="/dev/shm")
path = os.path.join(td, "clip.npy")
arr = np.lib.format.open_memmap(path, mode="w+", dtype=np.uint8, shape=(n_chunks * 512, 256, 256, 3))
rng = np.random.default_rng(0)
for c in range(n_chunks):
    arr[c * 512:(c + 1) * 512] = rng.integers(0, 256, (512, 256, 256, 3), dtype=np.uint8)
arr.flush(); del arr
gui_state.proj = None
cbas.encode_file(enc, path)  # warm-up
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cbas.encode_file(enc, path)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"encode_file: {dt * 1000:.1f} ms for {n_chunks} chunks = {dt / n_chunks * 1000:.2f} ms/chunk, {n_chunks * 512 / dt:.0f} frames/s")
# device-only time of the same chunks
x = torch.randint(0, 256, (512, 256, 256, 3), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n_chunks): enc.encode_u8(x)
torch.cuda.synchronize(); print(f"device only: {(time.perf_counter() - t0) / n_chunks * 1000:.2f} ms/chunk")
pr = cProfile.Profile(); pr.enable()
cbas.encode_file(enc, path)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
import shutil; shutil.rmtree(td, ignore_errors=True)
