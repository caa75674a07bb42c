import ctypes as C, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jellyfish_b200 import HashCounter, _lib
lib = _lib.load()
n_bases = 1_000_000_000
nbytes = lib.jfgpu_synth_fasta_bytes(n_bases)
text = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
got = C.c_uint64(0)
assert lib.jfgpu_synth_fasta_device(0, C.c_void_p(text.data_ptr()), nbytes + 256, n_bases, 7, C.byref(got), None) == 0
torch.cuda.synchronize()
world = 2
batch = 256 << 20
cap = int(batch / world * 1.25) + 65536
hc = HashCounter(8_000_000_000, 7, k=21, canonical=True, shard_index=0, n_shards=world, allow_regrow=False, max_batch_bytes=batch)
print(hc.info())
keys = torch.empty((world, cap), dtype=torch.int64, device="cuda")
counts = torch.zeros(world, dtype=torch.int64, device="cuda")
off = 0
stream = torch.cuda.current_stream().cuda_stream
for step in range(2):
    hc.clear()
    off = 0
    while off < got.value:
        ln = min(batch, got.value - off)
        counts.zero_()
        try:
            hc.extract_route(text.data_ptr() + off, ln, keys.data_ptr(), cap, counts.data_ptr(), begin=off == 0, end=off + ln >= got.value, stream=stream)
        except Exception as e:
            print("ERR", e, counts.tolist(), cap, hc.stats())
            raise
        c = counts.tolist()
        print(step, off, ln, c, cap)
        hc.insert_keys(keys[0].data_ptr(), c[0], stream=stream)
        off += ln
    print(hc.done())
