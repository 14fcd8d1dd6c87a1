"""Where does NCCL's log go?  One-rank communicator, three settings (tuning aid for bench.py's NCCL logging)."""
import os, subprocess, sys
code = r'''
import os, torch, torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
t = torch.ones(4, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
dist.destroy_process_group()
'''
for name, env in (("default", {}), ("file_stderr", {"NCCL_DEBUG_FILE": "/dev/stderr"}), ("file_path", {"NCCL_DEBUG_FILE": "/tmp/nccl_probe_%p.log"})):
    e = dict(os.environ, NCCL_DEBUG="INFO", NCCL_DEBUG_SUBSYS="INIT", **env)
    p = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True)
    files = [f for f in os.listdir("/tmp") if f.startswith("nccl_probe_")]
    nfile = sum(open("/tmp/" + f).read().count("NCCL INFO") for f in files)
    print(name, "stdout:", p.stdout.count("NCCL INFO"), "stderr:", p.stderr.count("NCCL INFO"), "file:", nfile, "rc", p.returncode)
    for line in (p.stdout + p.stderr).split("\n"):
        if "nranks" in line: print("   ", line[:160]); break
