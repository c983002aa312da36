"""Turns the reference's own normal_id_glm posterior fixture into committed test data.

The reference pins this path with one known-answer posterior test:
  data      src/test/unit/services/pathfinder/normal_glm_test.json      (N = 5000, K = 5)
  program   src/test/test-models/good/services/normal_glm.stan          (brms: centred X, normal_id_glm_lupdf,
            Intercept ~ N(0, 3), b ~ N(0, 3), sigma ~ N(1, 2))
  answers   src/test/unit/services/pathfinder/util.hpp:494-504          normal_glm_param_summary(): posterior
            means / SDs of (lp_approx__, lp__, b[1..5], Intercept, sigma, b_Intercept)
  bars      src/test/unit/services/pathfinder/normal_glm_test.cpp:130-135   |mean diff| < 0.01, |sd diff| < 0.1

/root/reference does not exist on the GPU box, so the data (re-serialised, values unchanged) and the expected
numbers are written here:  python tests/golden/make_normal_glm_fixture.py
"""
import json
import os
import re

REF = "/root/reference/src/test/unit/services/pathfinder"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    with open(os.path.join(REF, "normal_glm_test.json")) as f:
        d = json.load(f)
    assert d["N"] == 5000 and d["K"] == 5 and d["prior_only"] == 0
    with open(os.path.join(HERE, "normal_glm_data.json"), "w") as f:      # CmdStan JSON data format
        json.dump({"N": d["N"], "K": d["K"], "Y": d["Y"], "X": d["X"], "prior_only": 0}, f, separators=(",", ":"))
    src = open(os.path.join(REF, "util.hpp")).read()
    body = src[src.index("normal_glm_param_summary()"):]
    nums = lambda blk: [float(x) for x in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", blk)]
    mean = nums(body[body.index("mean_param_vals <<") + 18:body.index(";", body.index("mean_param_vals <<"))])
    sd = nums(body[body.index("sd_param_vals <<") + 16:body.index(";", body.index("sd_param_vals <<"))])
    names = ["lp_approx__", "lp__", "b.1", "b.2", "b.3", "b.4", "b.5", "Intercept", "sigma", "b_Intercept"]
    assert len(mean) == 10 and len(sd) == 10, (mean, sd)
    exp = {"source": "src/test/unit/services/pathfinder/util.hpp:494-504 (normal_glm_param_summary)",
           "bars": {"mean_abs": 0.01, "sd_abs": 0.1,
                    "source": "src/test/unit/services/pathfinder/normal_glm_test.cpp:130-135 (columns 2..9)"},
           "priors": {"prior_alpha_sd": 3.0, "prior_beta_sd": 3.0, "prior_sigma_loc": 1.0, "prior_sigma_scale": 2.0,
                      "source": "src/test/test-models/good/services/normal_glm.stan:30-32"},
           "names": names, "mean": mean, "sd": sd}
    with open(os.path.join(HERE, "normal_glm_expected.json"), "w") as f:
        json.dump(exp, f, indent=1)
    print("wrote normal_glm_data.json, normal_glm_expected.json", mean, sd)


if __name__ == "__main__":
    main()
