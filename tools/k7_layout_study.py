"""Bank-conflict study of the k = 7 histogram layout of long_kernel MODE_K7: why the write-out
goes through a host-built conflict-free schedule and a full row image instead of a direct gather.  Pure numpy."""
import numpy as np, sys
k=7
def rc(x,k):
    r=0
    for i in range(k):
        r=(r<<2)|((x&3)^3); x>>=2
    return r
canon=[x for x in range(4**k) if x<=rc(x,k)]
dim=len(canon)
# storage word index for rank j
def word_of(c):
    s=c
    if (s>>6)&3 >= 2: s=rc(c,k)
    assert (s>>6)&3 < 2
    hi=s&0xFF; lo=(s>>8)&0x3F
    return (hi<<6)|lo
src=np.array([word_of(c) for c in canon])
assert len(set(src))==dim and src.max()<8192
srcb=src%32; dstb=np.arange(dim)%32
# direct gather conflicts: 32 consecutive ranks
def wavefronts(groups):
    return sum(np.bincount(srcb[g],minlength=32).max() for g in groups)
direct=[np.arange(i,i+32) for i in range(0,dim,32)]
print("direct (consecutive ranks) LDS wavefronts:", wavefronts(direct), "ideal", dim//32)
# chunked: within each chunk of C ranks, how many groups needed = max over src banks multiplicity (dst banks uniform)
for C in (512,1024,2048,4096,8192):
    tot=0
    for i in range(0,dim,C):
        m=np.bincount(srcb[i:i+C],minlength=32).max()
        tot+=m
    print("chunk",C,"groups needed",tot,"ideal",dim//32)

# ---- XOR swizzles of the key layout: wavefronts of a DIRECT gather (lane l <- rank base + l), and groups a chunked
# colouring needs when the row image covers only C ranks at a time
base = np.array([((((c if (c >> 6) & 3 < 2 else rc(c, k)) & 0xFF) << 8) | ((((c if (c >> 6) & 3 < 2 else rc(c, k)) >> 8) & 0x3F) << 2))
                 for c in canon], dtype=np.int64)


def wf_direct(off):
    b = (off >> 2) % 32
    return int(sum(np.bincount(b[blk + 32 * e: blk + 32 * e + 32], minlength=32).max() for blk in range(0, dim, 128) for e in range(4)))


def chunk_groups(off, C):
    b = (off >> 2) % 32
    return int(sum(np.bincount(b[i:i + C], minlength=32).max() for i in range(0, dim, C)))


print("direct gather, no swizzle:", wf_direct(base), "wavefronts (ideal 256)")
for s in (3, 4, 5, 6, 7, 8):
    o = base ^ ((base >> s) & 0x7C)
    print(f"xor (off >> {s}) & 0x7C: direct {wf_direct(o)}, chunked colouring", {C: chunk_groups(o, C) for C in (512, 1024, 2048, 4096)})
best = []
for s1 in range(4, 10):
    for s2 in range(1, 8):
        for m2 in (0x04, 0x08, 0x10, 0x20, 0x40, 0x0C, 0x18, 0x30, 0x60):
            o = base ^ ((base >> s1) & 0x7C) ^ ((base >> s2) & m2)
            if len(set((o >> 2).tolist())) == dim:
                best.append((wf_direct(o), s1, s2, hex(m2)))
print("best two-term XOR swizzles (direct gather wavefronts, s1, s2, mask2):", sorted(best)[:5])
