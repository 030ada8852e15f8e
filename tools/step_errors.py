import sys, json
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import gpu_cases as g
from conftest import load_golden
from protein_redesign_b200 import synthetic as syn
out={}
for tag,cfg,sizes,seed in (("paper_n72",syn.PAPER,((12,60),(9,50)),3),("readme_n40",syn.README,((8,32),(6,27)),2),("paper_n128",syn.PAPER,((16,112),),4)):
    r=g.case_step(cfg,sizes,seed=seed,golden=load_golden(f"step_{tag}.npz"))
    out[tag]={k:float(f"{v[0]:.3g}") for k,v in r.items() if k in ("noise","seq","noise_vs_reference","seq_vs_reference")}
r=g.case_sample_T50(); out["T50"]={k:float(f"{v[0]:.3g}") for k,v in r.items()}
r=g.CASES["loss_paper_n72"](); out["loss"]={k:float(f"{v[0]:.3g}") for k,v in r.items() if k in ("loss","d_noise_pred","d_seq_pred","kernel_loss","kernel_d_seq_pred")}
print(json.dumps(out))
