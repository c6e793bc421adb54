"""Achieved HBM bandwidth of the bandwidth-bound launches of one train step (batch 16 @512x512) from the per-launch
CUDA-event table: algorithmic bytes (tensors read + written once, bf16 activations) / measured time, against the
measured copy bandwidth of MEASURED_PEAKS.json.  CUDA-event times carry ~5 us of event overhead per launch, so the small
tensors read low; the large ones are the statement.  usage: python scripts/elementwise_roofline.py <train_launches.csv>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, H, W = 16, 512, 512
ENC = {"layer1": (64, 4), "layer2": (128, 8), "layer3": (256, 16), "layer4": (512, 32)}
DEC = [(256, 16), (128, 8), (64, 4), (32, 2), (16, 1)]


def unit_elems(layer):
    """elements of the BatchNorm'd tensor behind a conv name, and whether the unit has a residual input"""
    if layer.startswith("encoder.conv1"):
        return 64 * (H // 2) * (W // 2) * N, False
    if layer.startswith("encoder.layer"):
        lay = layer.split(".")[1]
        c, d = ENC[lay]
        return c * (H // d) * (W // d) * N, ".conv2." in layer
    if layer.startswith("decoder.blocks."):
        b = int(layer.split(".")[2])
        c, d = DEC[b]
        return c * (H // d) * (W // d) * N, False
    raise KeyError(layer)


def main(path):
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    rows = list(csv.DictReader(open(path)))
    out = []
    for r in rows:
        k, layer, ms = r["kind"], r["layer"], float(r["ms"])
        try:
            if k == "bn_apply":
                e, res = unit_elems(layer)
                b = 2 * e * (2 + (1 if res else 0))
            elif k == "bn_bwd_reduce":
                e, res = unit_elems(layer)
                b = 2 * e * (2 + (1 if res else 0))
            elif k == "bn_bwd_apply":
                e, res = unit_elems(layer)
                b = 2 * e * (3 + (2 if res else 0))
            elif k == "maxpool":
                b = 2 * N * 64 * ((H // 2) * (W // 2) + (H // 4) * (W // 4)) + N * 64 * (H // 4) * (W // 4)
            elif k == "maxpool_bwd":
                b = 2 * N * 64 * (2 * (H // 2) * (W // 2) + (H // 4) * (W // 4)) + N * 64 * (H // 4) * (W // 4)
            elif k == "pack_input":
                b = N * H * W * 3 * 4 + N * H * (W + 8) * 4 * 2
            elif k == "head_bwd":
                b = N * H * W * 4 + N * H * W * 16 * 2
            else:
                continue
        except KeyError:
            continue
        out.append((k, layer, b / 1e6, ms * 1e3, b / ms / 1e6, b / ms / 1e6 / peak))
    print("kind,layer,algorithmic_MB,us,GBps,frac_of_measured_hbm_peak")
    for o in sorted(out, key=lambda o: -o[2]):
        print(f"{o[0]},{o[1]},{o[2]:.1f},{o[3]:.1f},{o[4]:.0f},{o[5]:.3f}")
    big = [o for o in out if o[2] >= 100]
    tb = sum(o[2] for o in big)
    tt = sum(o[3] for o in big)
    print(f"# launches with >= 100 MB: {len(big)}, {tb / 1e3:.2f} GB in {tt / 1e3:.3f} ms = {tb / tt * 1e3:.0f} GB/s "
          f"= {tb / tt * 1e3 / peak:.2f} of the measured {peak:.0f} GB/s")


if __name__ == "__main__":
    main(sys.argv[1])
