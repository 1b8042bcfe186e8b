"""SM clock / power while the forward runs back to back for a few seconds (nvidia-smi sampled every 20 ms)."""
import subprocess, sys, tempfile, time, torch
sys.path.insert(0, '/root/repo')
import bench
import speech_separation_b200 as V
dev = torch.device('cuda:0')
torch.manual_seed(42)
net = V.DPTNAVWavEncDec(**bench.MODEL_KW).eval().to(dev)
mix, s1, s2, e1, e2 = (t.to(dev) for t in bench.make_batch(32, 64000, 1234))
for _ in range(3): net(mix=mix, s1_embedding=e1, s2_embedding=e2)
torch.cuda.synchronize()
f = tempfile.NamedTemporaryFile("w+", suffix=".csv")
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,temperature.gpu,power.limit",
                      "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"], stdout=f)
time.sleep(0.3)
t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < 4.0:
    net(mix=mix, s1_embedding=e1, s2_embedding=e2); n += 1
    if n % 8 == 0: torch.cuda.synchronize()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
p.terminate(); p.wait(); f.flush(); f.seek(0)
rows = [l.split(",") for l in f.read().splitlines() if l.count(",") >= 5]
clk = sorted(float(r[0]) for r in rows); pw = sorted(float(r[1]) for r in rows)
cap = sum(1 for r in rows if r[2].strip().lower().startswith("active"))
print(f"{n} forwards in {dt:.2f} s = {dt/n*1e3:.2f} ms each; {len(rows)} samples: SM clock min/median/max {clk[0]:.0f}/{clk[len(clk)//2]:.0f}/{clk[-1]:.0f} MHz, "
      f"power median/max {pw[len(pw)//2]:.0f}/{pw[-1]:.0f} W (limit {rows[0][5].strip()} W), sw_power_cap active in {cap}/{len(rows)} samples, temp {rows[-1][4].strip()} C")
