"""Simulation of two ways to place the matches of a 32-token batch in the device inflate kernel (csrc/inflate.cu), on the real
token streams of sampled BGZF blocks, checked against zlib (analysis tool, CPU only):

    python tools/inflate_rounds_sim.py <file.bam> [n_blocks=25]

  serial : what the kernel does - literals and independent short matches in parallel, then every match that reads the output of
           its own batch or is longer than 16 bytes one at a time, whole warp per token, in stream order
  rounds : every pending short match whose source tokens are complete copies in the same round (source tokens found by a binary
           search of the batch's prefix sums); the lowest pending long match is copied by the whole warp between rounds
C2: 8.07 -> 3.43 serial iterations per batch. The loop sits in the ~20 % of the kernel's instructions that are not lane 0's
Huffman decode (profiles/r1_summary.md), so the rounds are worth at most ~10 % of the kernel time: second priority after the
instruction count of the decode loop itself.
"""
import sys,collections,random,struct,zlib
import os
sys.path.insert(0,os.path.dirname(os.path.abspath(__file__)))
import deflate_stats as D
# token stream per BGZF block
def tokens_of(payload):
    b=D.Bits(payload); toks=[]
    while True:
        final,typ=b.get(1),b.get(2)
        assert typ==2
        hlit,hdist,hclen=b.get(5)+257,b.get(5)+1,b.get(4)+4
        cl=[0]*19
        for i in range(hclen): cl[D.CL_ORDER[i]]=b.get(3)
        clt=D.make_decoder(cl); lens=[]
        while len(lens)<hlit+hdist:
            s,_=D.decode_sym(b,clt)
            if s<16: lens.append(s)
            elif s==16: lens+=[lens[-1]]*(3+b.get(2))
            elif s==17: lens+=[0]*(3+b.get(3))
            else: lens+=[0]*(11+b.get(7))
        lt,dt=D.make_decoder(lens[:hlit]),D.make_decoder(lens[hlit:])
        while True:
            s,l=D.decode_sym(b,lt)
            if s==256: break
            if s<256: toks.append((1,s,0)); continue
            i=s-257; ln=D.LEN_BASE[i]+b.get(D.LEN_EXTRA[i]); ds,dl=D.decode_sym(b,dt); dist=D.DIST_BASE[ds]+b.get(D.DIST_EXTRA[ds])
            toks.append((0,ln,dist))
        if final: break
    return toks
COOP=16
def place(toks, scheme, st):
    out=bytearray(); 
    for bstart in range(0,len(toks),32):
        batch=toks[bstart:bstart+32]; pos=len(out)
        offs=[];o=0
        for lit,a,b_ in batch: offs.append(o); o+= 1 if lit else a
        total=o; out+=bytes(total)
        n=[1 if t[0] else t[1] for t in batch]
        done=[False]*len(batch); pending_short=[];pending_long=[]
        # phase A
        for i,(lit,a,dist) in enumerate(batch):
            if lit: out[pos+offs[i]]=a; done[i]=True
            else:
                coop = a>COOP or dist<offs[i]+a
                if not coop:
                    for k in range(a): out[pos+offs[i]+k]=out[pos+offs[i]+k-dist]
                    done[i]=True
        st['batches']+=1
        pend=[i for i in range(len(batch)) if not done[i]]
        if scheme=='serial':
            for i in pend:
                a,dist=batch[i][1],batch[i][2]
                for k in range(a): out[pos+offs[i]+k]=out[pos+offs[i]+k-dist]
                st['iters']+=1
        else:
            import bisect
            dep={}
            for i in pend:
                a,dist=batch[i][1],batch[i][2]
                lo=max(0,offs[i]-dist); hi=min(offs[i],offs[i]-dist+a)  # [lo,hi) inside batch output produced by earlier tokens
                if hi>lo:
                    jl=bisect.bisect_right(offs,lo)-1; jh=bisect.bisect_right(offs,hi-1)-1
                    dep[i]=(jl,jh)
                else: dep[i]=None
            while pend:
                st['iters']+=1
                snapshot=list(done)
                ready=[i for i in pend if batch[i][1]<=COOP and (dep[i] is None or all(snapshot[j] for j in range(dep[i][0],dep[i][1]+1)))]
                # parallel semantics: compute from a frozen copy of out for cross-lane reads
                frozen=bytes(out)
                for i in ready:
                    a,dist=batch[i][1],batch[i][2]
                    for k in range(a):
                        srcpos=pos+offs[i]+k-dist
                        out[pos+offs[i]+k]= out[srcpos] if srcpos>=pos+offs[i] else frozen[srcpos]
                    done[i]=True
                pend=[i for i in pend if not done[i]]
                if pend and batch[pend[0]][1]>COOP:
                    i=pend[0]; a,dist=batch[i][1],batch[i][2]
                    for k in range(a): out[pos+offs[i]+k]=out[pos+offs[i]+k-dist]
                    done[i]=True; pend=pend[1:]; st['long_steps']+=1
                elif not ready:
                    raise SystemExit('stuck')
    return bytes(out)
raw=open(sys.argv[1],'rb').read()
offs=[];o=0
while o<len(raw):
    bs=struct.unpack_from("<H",raw,o+16)[0]+1; offs.append((o,bs)); o+=bs
rng=random.Random(2)
S1=collections.Counter();S2=collections.Counter()
for o,bs in rng.sample(offs[1:-1],int(sys.argv[2]) if len(sys.argv)>2 else 25):
    payload=raw[o+18:o+bs-8]; want=zlib.decompress(payload,-15)
    toks=tokens_of(payload)
    assert place(toks,'serial',S1)==want
    assert place(toks,'rounds',S2)==want
print('serial: iterations per batch %.2f'%(S1['iters']/S1['batches']))
print('rounds: iterations per batch %.2f (of which long-token steps %.2f)'%(S2['iters']/S2['batches'],S2['long_steps']/S2['batches']))
