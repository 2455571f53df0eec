import sys, torch
sys.path.insert(0, '.')
import mmnas_b200
from mmnas_b200 import kernels as K
torch.manual_seed(0)
dev='cuda'
M,H,Fd = 400,512,2048
x = torch.randn(M,H,device=dev); W1 = torch.randn(Fd,H,device=dev)/H**0.5; b1=torch.randn(Fd,device=dev)*0.1
W2 = torch.randn(H,Fd,device=dev)/Fd**0.5
db = torch.randn(M,H,device=dev)
x16,W116,W216,db16 = x.bfloat16(),W1.bfloat16(),W2.bfloat16(),db.bfloat16()
h = torch.empty(M,Fd,device=dev,dtype=torch.bfloat16)
K.gemm_bf16(M,Fd,H,x16,H,0,W116,H,0,h,Fd,bias=b1,relu=True)
href = torch.relu(x16.float()@W116.float().t()+b1)
print('h err', (h.float()-href).abs().max().item(), 'mask mismatch', ((h>0)!=(href.bfloat16()>0)).sum().item())
dh = torch.empty(M,Fd,device=dev,dtype=torch.bfloat16)
K.gemm_bf16(M,Fd,H,db16,H,0,W216,Fd,1,dh,Fd,aux=h,ld_aux=Fd,aux_scale=1.0)
dhref = (db16.float()@W216.float())*(h>0)
e = (dh.float()-dhref).abs()
print('dh err max', e.max().item(), 'ref max', dhref.abs().max().item())
colerr = e.max(0).values
print('bad cols', (colerr>0.05).nonzero().flatten()[:40].tolist(), 'n', (colerr>0.05).sum().item())
rowerr = e.max(1).values
print('bad rows', (rowerr>0.05).nonzero().flatten()[:40].tolist(), 'n', (rowerr>0.05).sum().item())
db1 = torch.empty(Fd,device=dev); K.colsum(dh,M,Fd,Fd,db1)
print('colsum err vs own dh', (db1-dh.float().sum(0)).abs().max().item(), 'vs ref', (db1-dhref.sum(0)).abs().max().item(), dhref.sum(0).abs().max().item())
dW1 = torch.zeros(Fd,H,device=dev)
K.gemm_bf16(Fd,H,M,dh,Fd,1,x16,H,1,dW1,H,split_k=1)
ref = dh.float().t()@x16.float()
print('dW1 err vs own dh', (dW1-ref).abs().max().item(), ref.abs().max().item())
ee=(dW1-ref).abs(); print('bad dW1 rows', (ee.max(1).values>0.05).sum().item(), 'cols', (ee.max(0).values>0.05).sum().item())
for sk in (2,3):
    dW1.zero_(); K.gemm_bf16(Fd,H,M,dh,Fd,1,x16,H,1,dW1,H,split_k=sk); print('split',sk,(dW1-ref).abs().max().item())
