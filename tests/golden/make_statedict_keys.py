"""Record the reference's state_dict key names and shapes (needs /root/reference; build container only)."""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, make_opt  # noqa: E402
PoseNDF, _, _ = import_reference()
out = {}
for use_enc in (True, False):
    net = PoseNDF(make_opt(dict(use_enc=use_enc, enc_act="lrelu", enc_beta=100, df_act="lrelu", df_beta=100)))
    out["enc" if use_enc else "noenc"] = [[k, list(v.shape)] for k, v in net.state_dict().items()]
json.dump(out, open(os.path.join(HERE, "statedict_keys.json"), "w"), indent=0)
print({k: len(v) for k, v in out.items()})
