"""Development aid: repeat run_device with concurrent embed / chunk streams; dump all Python stacks if it stalls."""
import faulthandler
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.enable()
import torch

from tools.workload import make_session_cfg
from whisper_nemo_b200 import ClusteringDiarizer, checkpoint

domain, seconds, reps = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
weights = checkpoint.seeded()
if domain.startswith("batch"):  # batchN: N recordings in one manifest (general YAML)
    from tools import workload as synth
    from whisper_nemo_b200 import config

    root = tempfile.mkdtemp()
    entries = []
    for i in range(int(domain[5:])):
        w, r, _, _ = synth.make_session(root, f"rec{i}", seconds, 2 + i % 3, seed=40 + i)
        entries.append({"audio_filepath": w, "rttm_filepath": r})
    cfg = config.load_config("general")
    synth.write_manifest(os.path.join(root, "m.json"), entries)
    cfg.diarizer.manifest_filepath, cfg.diarizer.out_dir, cfg.diarizer.oracle_vad = os.path.join(root, "m.json"), root, True
else:
    cfg, _, _ = make_session_cfg(tempfile.mkdtemp(), domain, seconds, 3, seed=3)
diar = ClusteringDiarizer(cfg=cfg, speaker_model=weights)
diar._prepare()
wav = diar._wav_host.to(dev)
for i in range(reps):
    faulthandler.dump_traceback_later(25, exit=True)
    t0 = time.time()
    diar.run_device(wav_dev=wav, timers=False)
    torch.cuda.synchronize()
    faulthandler.cancel_dump_traceback_later()
    print(f"rep {i} ok {time.time() - t0:.3f}s", flush=True)
print("DONE", flush=True)
