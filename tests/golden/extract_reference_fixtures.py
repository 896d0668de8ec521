"""Transcribes the reference's own stored known-answer outputs into JSON fixtures.

Run in the build container (needs /root/reference); the resulting files are
committed so that tests never read /root/reference at run time.

Sources (verbatim cell outputs, not recomputed):
  doc/Gpx_Tutorial.ipynb:165-167   5-point kriging theta / variance / likelihood
  doc/Gpx_Tutorial.ipynb:420-421   full serde-JSON of a trained Linear+Matern52 model
"""
import json
import re
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).parent


def main():
    nb = json.loads((REF / "doc" / "Gpx_Tutorial.ipynb").read_text())
    texts = []
    for cell in nb["cells"]:
        for out in cell.get("outputs", []):
            if "text" in out:
                texts.append("".join(out["text"]))
    # --- 5-point kriging printed values
    t = next(s for s in texts if "Optimal theta" in s)
    five = {
        "source": "doc/Gpx_Tutorial.ipynb:165-167",
        "xt": [0.0, 1.0, 2.0, 3.0, 4.0],
        "yt": [0.0, 1.0, 1.5, 0.9, 1.0],
        "theta": float(re.search(r"Optimal theta = \[([0-9.eE+-]+)\]", t).group(1)),
        "variance": float(re.search(r"GP variance = ([0-9.eE+-]+)", t).group(1)),
        "likelihood": float(re.search(r"Reduced likelihood = ([0-9.eE+-]+)", t).group(1)),
    }
    (OUT / "gpx_tutorial_kriging5.json").write_text(json.dumps(five, indent=1))
    # --- full model JSON
    t = next(s for s in texts if "Gpx stringified JSON serialization" in s)
    js = t[t.index("{", t.index("Gpx stringified JSON serialization")):].strip()
    # the notebook output is cut mid-way by the stream size limit: keep the first expert only
    start = js.index('"experts":[') + len('"experts":[')
    depth, i = 0, start
    while True:
        c = js[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    expert = json.loads(js[start:i + 1])
    expert["source"] = "doc/Gpx_Tutorial.ipynb:420-421 (experts[0])"
    (OUT / "gpx_tutorial_linear_matern52.json").write_text(json.dumps(expert, indent=1))
    # --- mixture-level blocks of the same stored model (gmx of the one-cluster GMM, the params with the
    #     regression / correlation spec sets the expert was SELECTED from by cross-validation), plus the
    #     `Gpx string` line printed next to it
    full = json.loads(js)
    mix = {k: full[k] for k in ("recombination", "gmx", "gp_type", "training_data", "params")}
    mix["display"] = re.search(r"Gpx string: (.*)", t).group(1).strip()
    mix["selected_expert"] = full["experts"][0]["type_fullgp"]
    mix["source"] = "doc/Gpx_Tutorial.ipynb:420-421 (mixture-level blocks)"
    (OUT / "gpx_tutorial_mixture.json").write_text(json.dumps(mix, indent=1))
    print("wrote", [p.name for p in OUT.glob("gpx_tutorial_*.json")])


if __name__ == "__main__":
    main()
