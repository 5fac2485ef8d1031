"""CPU simulation (numpy, no GPU) of the number of 32-byte L2 sector reductions the fused backward issues per coordinate and
level, for a batch drawn like the bench draws it (2^19 of the 352x352x6x15 voxels, no replacement) under different in-batch
orders and register-merge rules.  The instruction-level lane mapping is the kernel's: one instruction = 8 consecutive rows x the
two axis-0 halves of one (y,z,t) corner combination; lanes that hit one sector at different addresses are one sector
operation, lanes that hit the same address are separate operations unless merged in registers first.
Output (sector operations per coordinate, per level and summed) backs DESIGN.md section 3: line order without merge 154.8,
the shipped line-run merge on 8 levels 123.6 (measured by ncu: 143.2, a constant 1.16x above the model in every
configuration incl. round 1's i.i.d. order), ideal in-group merge 122.3, 2-D tile orders with a per-cell merge 121.2."""
import numpy as np
rng=np.random.default_rng(0)
shape=(352,352,6,15)
N=int(np.prod(shape)); n=1<<19
idx=np.sort(rng.choice(N,n,replace=False))
# C-order index -> coords (x=axis0 slowest in C order!).  locality order: axis-0 fastest
i3=idx%15; r=idx//15; i2=r%6; r//=6; i1=r%352; i0=r//352
def order_line():
    key=((i3.astype(np.int64)*6+i2)*352+i1)*352+i0
    return np.argsort(key,kind='stable')
def order_tile(tx,ty):
    # tiles of tx x ty in (axis0, axis1), inside a tile axis-0 fastest
    key=(((i3.astype(np.int64)*6+i2)*(352//ty+1)+i1//ty)*(352//tx+1)+i0//tx)
    key=(key*ty+i1%ty)*tx+i0%tx
    return np.argsort(key,kind='stable')
res=[16,22,31,43,61,86,120,168,236,330,462,647,907,1269,1777,2489]
rows=[65536,234256]+[524288]*14
P=[1,2654435761,805459861,3674653429]
def count(order,merge='line',take=1<<16):
    o=order[:take]
    X=np.stack([i0[o]/351.0,i1[o]/351.0,i2[o]/5.0,i3[o]/14.0],1).astype(np.float32)
    tot=[]
    for l,(rs,T) in enumerate(zip(res,rows)):
        xs=(X*np.float32(rs)); ci=xs.astype(np.int64)
        ops=0
        nb=take//8
        c=ci.reshape(nb,8,4)
        for combo in range(8):
            base=np.zeros((nb,8),np.uint64)
            for d in (1,2,3):
                up=(combo>>(d-1))&1
                base^=((c[:,:,d]+up).astype(np.uint64)*np.uint64(P[d]))&np.uint64(0xffffffff)
            for nothing in (0,):
                lo=(base^c[:,:,0].astype(np.uint64))&np.uint64(0xffffffff)
                hi=(base^(c[:,:,0]+1).astype(np.uint64))&np.uint64(0xffffffff)
                addr=np.concatenate([lo,hi],1)%np.uint64(T)   # (nb,16)
                # key for merging: samples with identical (y,z,t) coords + same x cell -> same address AND mergeable
                if merge=='none':
                    # every lane own op unless same sector different address
                    a=np.sort(addr,1)
                    sect=a//np.uint64(4)
                    # count: distinct sectors + (duplicates of same address)
                    dup_addr=(a[:,1:]==a[:,:-1]).sum()
                    dist_sect=(np.diff(sect.astype(np.int64),axis=1)!=0).sum()+nb
                    ops+=dist_sect+dup_addr
                else:
                    # ideal merge of all same-address lanes within the 16-lane group (upper bound of any register merge)
                    a=np.sort(addr,1)
                    sect=a//np.uint64(4)
                    dist_sect=(np.diff(sect.astype(np.int64),axis=1)!=0).sum()+nb
                    ops+=dist_sect
        tot.append(ops/take)
    return tot
ol=order_line()
for name,o,m in (('line,no-merge',ol,'none'),('line,ideal-merge',ol,'ideal')):
    t=count(o,m); print(name,' '.join(f'{v:5.1f}' for v in t),' sum',round(sum(t),1))
for tx,ty in ((8,8),(16,16),(32,8),(16,4),(64,4)):
    o=order_tile(tx,ty)
    t=count(o,'ideal'); print(f'tile{tx}x{ty},ideal',' '.join(f'{v:5.1f}' for v in t),' sum',round(sum(t),1))

def count_kernel(order, take=1<<16, merge_levels=8):
    o=order[:take]
    X=np.stack([i0[o]/351.0,i1[o]/351.0,i2[o]/5.0,i3[o]/14.0],1).astype(np.float32)
    line_id=((i3[o].astype(np.int64)*6+i2[o])*352+i1[o]).reshape(-1,8)
    tot=[]
    nb=take//8
    for l,(rs,T) in enumerate(zip(res,rows)):
        xs=(X*np.float32(rs)); ci=xs.astype(np.int64)
        c=ci.reshape(nb,8,4)
        # which lanes issue (after merging)?  lower-half lane of row g: kk=cell ; upper: kk=cell+1
        cell=c[:,:,0]
        same_line=np.zeros((nb,8),bool); same_line[:,1:]=line_id[:,1:]==line_id[:,:-1]
        issue_lo=np.ones((nb,8),bool); issue_hi=np.ones((nb,8),bool)
        if l<merge_levels:
            cont=np.zeros((nb,8),bool); cont[:,1:]=same_line[:,1:]&(cell[:,1:]==cell[:,:-1])   # row continues previous row's run
            issue_lo=~cont; issue_hi=~cont                       # heads only
            # link: lower run starting at row g absorbs upper run ending at g-1 if same line and cell[g-1]+1==cell[g]
            link=np.zeros((nb,8),bool); link[:,1:]=same_line[:,1:]&(cell[:,:-1]+1==cell[:,1:])
            # the upper run that ends at row g-1: its head is not issued. find head of run containing g-1
            head_idx=np.zeros((nb,8),np.int64)
            for g in range(8):
                head_idx[:,g]=np.where(cont[:,g], head_idx[:,g-1] if g>0 else 0, g)
            for g in range(1,8):
                rows_sel=np.nonzero(link[:,g])[0]
                issue_hi[rows_sel, head_idx[rows_sel,g-1]]=False
        ops=0
        for combo in range(8):
            base=np.zeros((nb,8),np.uint64)
            for d in (1,2,3):
                up=(combo>>(d-1))&1
                base^=((c[:,:,d]+up).astype(np.uint64)*np.uint64(P[d]))&np.uint64(0xffffffff)
            lo=((base^cell.astype(np.uint64))&np.uint64(0xffffffff))%np.uint64(T)
            hi=((base^(cell+1).astype(np.uint64))&np.uint64(0xffffffff))%np.uint64(T)
            addr=np.concatenate([lo,hi],1).astype(np.int64)
            issued=np.concatenate([issue_lo,issue_hi],1)
            addr=np.where(issued,addr,-np.arange(1,17)[None,:]*4)   # unique negative sectors for non-issued, subtract later
            a=np.sort(addr,1)
            sect=a//4
            dist=(np.diff(sect,axis=1)!=0).sum()+nb
            dup=((a[:,1:]==a[:,:-1])&(a[:,1:]>=0)).sum()
            ops+=dist+dup-(~issued).sum()
        tot.append(ops/take)
    return tot
t=count_kernel(ol); print('kernel-sim line merge8',' '.join(f'{v:5.1f}' for v in t),' sum',round(sum(t),1))
t=count_kernel(ol,merge_levels=16); print('kernel-sim line merge16',' '.join(f'{v:5.1f}' for v in t),' sum',round(sum(t),1))

def count_kernel_cellmerge(order, take=1<<16, merge_levels=8):
    """merge keyed on (x cell, y cell) with identical (z,t): per-cell register merge (weights carried per y corner)"""
    o=order[:take]
    X=np.stack([i0[o]/351.0,i1[o]/351.0,i2[o]/5.0,i3[o]/14.0],1).astype(np.float32)
    zt=(i3[o].astype(np.int64)*6+i2[o]).reshape(-1,8)
    tot=[]
    nb=take//8
    for l,(rs,T) in enumerate(zip(res,rows)):
        xs=(X*np.float32(rs)); ci=xs.astype(np.int64)
        c=ci.reshape(nb,8,4)
        cell=c[:,:,0]; celly=c[:,:,1]
        same=np.zeros((nb,8),bool); same[:,1:]=(zt[:,1:]==zt[:,:-1])&(celly[:,1:]==celly[:,:-1])
        issue_lo=np.ones((nb,8),bool); issue_hi=np.ones((nb,8),bool)
        if l<merge_levels:
            cont=np.zeros((nb,8),bool); cont[:,1:]=same[:,1:]&(cell[:,1:]==cell[:,:-1])
            issue_lo=~cont; issue_hi=~cont
            link=np.zeros((nb,8),bool); link[:,1:]=same[:,1:]&(cell[:,:-1]+1==cell[:,1:])
            head_idx=np.zeros((nb,8),np.int64)
            for g in range(8):
                head_idx[:,g]=np.where(cont[:,g], head_idx[:,g-1] if g>0 else 0, g)
            for g in range(1,8):
                rows_sel=np.nonzero(link[:,g])[0]
                issue_hi[rows_sel, head_idx[rows_sel,g-1]]=False
        ops=0
        for combo in range(8):
            base=np.zeros((nb,8),np.uint64)
            for d in (1,2,3):
                up=(combo>>(d-1))&1
                base^=((c[:,:,d]+up).astype(np.uint64)*np.uint64(P[d]))&np.uint64(0xffffffff)
            lo=((base^cell.astype(np.uint64))&np.uint64(0xffffffff))%np.uint64(T)
            hi=((base^(cell+1).astype(np.uint64))&np.uint64(0xffffffff))%np.uint64(T)
            addr=np.concatenate([lo,hi],1).astype(np.int64)
            issued=np.concatenate([issue_lo,issue_hi],1)
            addr=np.where(issued,addr,-np.arange(1,17)[None,:]*4)
            a=np.sort(addr,1); sect=a//4
            dist=(np.diff(sect,axis=1)!=0).sum()+nb
            dup=((a[:,1:]==a[:,:-1])&(a[:,1:]>=0)).sum()
            ops+=dist+dup-(~issued).sum()
        tot.append(ops/take)
    return tot
for tx,ty in ((8,8),(16,4),(4,16),(2,32),(1,64),(32,2)):
    o=order_tile(tx,ty)
    t=count_kernel_cellmerge(o); print(f'cellmerge tile{tx}x{ty}',' '.join(f'{v:5.1f}' for v in t),' sum',round(sum(t),1))
t=count_kernel_cellmerge(ol); print('cellmerge line',' '.join(f'{v:5.1f}' for v in t),' sum',round(sum(t),1))
