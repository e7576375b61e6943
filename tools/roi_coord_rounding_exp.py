import sys, torch, torchvision, numpy as np
sys.path.insert(0,'/root/repo')
from livecell_instance_segmentation_b200 import ops, synth
dev='cuda:0'
g=torch.Generator(device=dev).manual_seed(1)
for (H,W,K,mode) in ((130,176,4096,"anchor"),(64,64,128,"anchor")):
    feat=torch.randn((1,256,H,W),generator=g,device=dev)
    fn=feat.contiguous(memory_format=torch.channels_last)
    rois=torch.from_numpy(synth.make_rois(K,5,img_h=H*4,img_w=W*4,mode=mode)).to(dev)
    tv=torchvision.ops.roi_align(feat,rois,(7,7),0.25,2,False)
    tv_cpu=torchvision.ops.roi_align(feat.cpu(),rois.cpu(),(7,7),0.25,2,False).to(dev)
    tv64=torchvision.ops.roi_align(feat.cpu().double(),rois.cpu().double(),(7,7),0.25,2,False).to(dev)
    a=ops.roi_align_fwd([fn],[0.25],rois,None,(7,7),2,False,cpu_coords=False)
    b=ops.roi_align_fwd([fn],[0.25],rois,None,(7,7),2,False,cpu_coords=True)
    sc=float(tv.abs().max())
    f=lambda x,y: float((x-y).abs().max())/sc
    print(f"fwd {H}x{W} K={K}: ours(cuda coords) vs TV-CUDA {f(a,tv):.2e} | ours(cpu coords) vs TV-CUDA {f(b,tv):.2e} | ours(cpu) vs TV-CPU {f(b,tv_cpu):.2e} | ours(cuda) vs TV-CPU {f(a,tv_cpu):.2e} | TV-CUDA vs TV-CPU {f(tv,tv_cpu):.2e} | vs fp64: TVcuda {f(tv,tv64):.2e} TVcpu {f(tv_cpu,tv64):.2e} ours_cuda {f(a,tv64):.2e} ours_cpu {f(b,tv64):.2e}")
    gout=torch.randn(tv.shape,generator=g,device=dev)
    fr=feat.clone().requires_grad_(True); torchvision.ops.roi_align(fr,rois,(7,7),0.25,2,False).backward(gout); gtv=fr.grad
    frc=feat.cpu().clone().requires_grad_(True); torchvision.ops.roi_align(frc,rois.cpu(),(7,7),0.25,2,False).backward(gout.cpu()); gtvc=frc.grad.to(dev)
    outs=[]
    for cc in (False,True):
        gin=torch.empty_like(fn); ops.roi_align_bwd(gout,[gin],[0.25],rois,None,2,False,zero_grad=True,cpu_coords=cc); outs.append(gin)
    sc=float(gtv.abs().max()); 
    print(f"bwd: ours(cuda) vs TV-CUDA {f(outs[0],gtv):.2e} | ours(cpu) vs TV-CUDA {f(outs[1],gtv):.2e} | ours(cpu) vs TV-CPU {f(outs[1],gtvc):.2e} | TV-CUDA vs TV-CPU {f(gtv,gtvc):.2e}")
