"""Live oracle-vs-reference comparison at full network size (needs /root/reference; run in its own
process because the reference's `nemo` package and this repo's mirror share the import name).
Prints one JSON line: mismatching activation codes / scales / logit codes."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [HERE, os.path.join(ROOT, "oracle")]
sys.path.append(os.path.join(ROOT, "q-asr_b200", "qasr"))     # topology / params only (no `nemo` mirror)

from ref_loader import RefModel, load_reference  # noqa: E402
import params  # noqa: E402
import topology  # noqa: E402
import qasr_oracle as O  # noqa: E402
from spec_from_modules import spec_from_model  # noqa: E402

torch.set_grad_enabled(False)


def main(topo, B, T, wbit, abit, mode, literal):
    ref = load_reference()
    Q, qm = ref.quant_modules, ref.quantize_model
    O.LITERAL_NOISE = bool(literal)
    torch.manual_seed(0)
    m = RefModel(ref, topology.TOPOLOGIES[topo]())
    m.eval()
    params.reestimate_bn_(m.encoder)
    m.set_quant_bit(wbit, "weight")
    m.set_quant_bit(abit, "act")
    m.encoder.bn_folding()
    spec = spec_from_model(m.encoder, m.decoder)
    if mode != "dynamic":
        qm.calibrate(m)
        O.set_mode(spec, True)
        for i in range(2):
            cx, cl = params.synthetic_features(B, T, seed=100 + i, ragged=(i % 2 == 1))
            e, l, s = m.encoder(cx, cl)
            m.decoder(e, s)
            O.model_forward(spec, cx.numpy(), cl.numpy())
    qm.evaluate(m)
    O.set_mode(spec, False, dynamic=(mode == "dynamic"))
    qm.set_dynamic(m, mode == "dynamic")
    got = {}

    def hook(n):
        def f(mod, inp, out):
            got.setdefault(n, []).append((torch.round(out[0] / out[1].reshape(1, -1, 1)).numpy(),
                                          out[1].numpy().reshape(-1)))
        return f
    for n, mod in m.named_modules():
        if type(mod) is Q.QuantAct and not n.startswith("encoder.encoder_layers"):
            mod.register_forward_hook(hook(n.replace("encoder.encoder.", "encoder.")))
    x, lens = params.synthetic_features(B, T, seed=5, ragged=True)
    e, l, s = m.encoder(x, lens)
    m.decoder(e, s)
    trace = []
    out = O.model_forward(spec, x.numpy(), lens.numpy(), trace)
    code_mism = total = scale_mism = 0
    seen = {}
    for rec in trace:
        name = rec["name"]
        i = seen.get(name, 0)
        seen[name] = i + 1
        if name.endswith("res_act"):
            rc, rs = got[name][i]
            c, sc = rec["codes"], rec["scale"]
        else:
            rc, rs = got["decoder.act" if name == "decoder" else name + ".act"][i]
            c, sc = rec["in_codes"], rec["in_scale"]
        code_mism += int((rc != c).sum())
        total += c.size
        scale_mism += int(np.float32(sc) != rs.reshape(-1)[0])
    ref_codes = torch.round(m.decoder.last_logits.double() / m.decoder.last_logit_sf.double()).numpy().astype(np.int64)
    print(json.dumps(dict(code_mismatches=code_mism, codes=total, scale_mismatches=scale_mism,
                          logit_code_mismatches=int((ref_codes != out["logit_codes"]).sum()),
                          logit_codes=int(ref_codes.size))))


if __name__ == "__main__":
    a = sys.argv[1:]
    main(a[0], int(a[1]), int(a[2]), int(a[3]), int(a[4]), a[5], int(a[6]))
