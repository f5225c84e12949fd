"""GPU debugging aid: runs the offline engine without scratch recycling and compares EVERY intermediate tensor
with the CPU oracle's taps.  Writes gpurun_out/layer_report.txt.   usage: python tests/layer_report.py [B T]"""
import os
import sys

os.environ["NUNET_DEBUG_KNOBS"] = "1"
os.environ["NUNET_NO_RECYCLE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from nunet_b200.engine import NunetEngine
from nunet_b200.synth import synth_clips
from nunet_b200.weights import load_default_weights, pack_blob
from oracle.nunet_oracle import Oracle


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    w = load_default_weights()
    wav = synth_clips(B, 512 + 256 * (T - 1))
    lines = []
    for mode in ("causal_avg32", "frame_div32"):
        o = Oracle(w, ctfa_mode=mode)
        mags, _ = o.stft(torch.from_numpy(wav))
        mag = mags[:, :, 1:].contiguous()
        taps = {}
        with torch.no_grad():
            ref = o.net(mag[..., None], taps=taps).squeeze(-1)
        eng = NunetEngine(pack_blob(w), max_frames=B * T, ctfa_mode=mode)
        out = eng.forward_mag(mag.cuda()).cpu()
        torch.cuda.synchronize()
        lines.append(f"== mode {mode}  B={B} T={T}  launches={eng.last_launch_count}")
        lines.append(f"{'tensor':28s} {'max|ref|':>10s} {'max|diff|':>10s}  first-bad(b,t,f,c)")
        for name, t in taps.items():
            try:
                got = eng.debug_read(name)
            except Exception as e:  # noqa: BLE001
                lines.append(f"{name:28s} -- {e}")
                continue
            r = t.numpy().reshape(-1)
            if got.size != r.size:
                lines.append(f"{name:28s} SIZE {got.size} != {r.size}")
                continue
            d = np.abs(got - r)
            bad = ""
            if d.max() > 1e-3 * max(1.0, np.abs(r).max()):
                idx = np.unravel_index(int(np.argmax(d > 1e-3 * max(1.0, np.abs(r).max()))), t.shape)
                bad = f"{idx} got {got.reshape(t.shape)[idx]:.5f} ref {t.numpy()[idx]:.5f}"
            lines.append(f"{name:28s} {np.abs(r).max():10.4f} {d.max():10.3e}  {bad}")
        d = (out - ref).abs()
        lines.append(f"{'model_out':28s} {float(ref.abs().max()):10.4f} {float(d.max()):10.3e}")
        eng.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "layer_report.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
