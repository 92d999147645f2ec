import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import phoskintime_b200 as pk
eng = pk.get_engine(0)
T14 = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
B = 1 << 18
gen = torch.Generator(device="cuda").manual_seed(11)
p = torch.rand((B, 14), generator=gen, device="cuda", dtype=torch.float64) * 2.95 + 0.05
t = torch.from_numpy(T14).cuda()
y0 = torch.rand((B, 7), generator=gen, device="cuda", dtype=torch.float64) + 0.1
for method in ("ros6l", "ros5l"):
    r = eng.solve_local_batch("succmod", p, y0, 5, t, want=("sol",), method=method)
    sol = r["sol"]; k = 7
    t2 = t[k:].clone()
    r2 = eng.solve_local_batch("succmod", p, sol[:, k].contiguous(), 5, t2, want=("sol",), method=method)
    tight = eng.solve_local_batch("succmod", p, y0, 5, t, want=("sol",), rtol=1e-10, atol=1e-14, method="ros5l")["sol"]
    d = (r2["sol"] - sol[:, k:]).abs() / (1e-6 * sol[:, k:].abs() + 1e-9)
    i = int(d.flatten().argmax()); b, kk, s = np.unravel_index(i, d.shape)
    print(method, "max", float(d.max()), "at sys", b, "time idx", kk, "state", s, "restart", float(r2["sol"][b,kk,s]), "orig", float(sol[b,k+kk,s]), "tight", float(tight[b,k+kk,s]),
          "nsteps orig", int(r["nsteps"][b]), "restart", int(r2["nsteps"][b]), "nrej", int(r2["nrej"][b]))
    e1 = ((sol - tight).abs() / (1e-6 * tight.abs() + 1e-9)).max(); 
    print("   orig vs tight max", float(e1), " count d>1:", int((d.amax(dim=(1,2)) > 1).sum()))
    print("   params", p[b].cpu().numpy().round(3), "y at restart", sol[b,k].cpu().numpy())
    print("   restart traj state", s, r2["sol"][b,:,s].cpu().numpy(), "\n   orig", sol[b,k:,s].cpu().numpy(), "\n   tight", tight[b,k:,s].cpu().numpy())
