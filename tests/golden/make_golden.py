"""Extract the golden vectors recoverable from the reference's notebooks (printed OUTPUTS only)
into tests/golden/reference_golden.json.  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Sources:
  D1/D2  Baby exact belief chain        notebooks/Belief_and_Particle_Filter_Options.ipynb, Display_Tree.ipynb
  D3     printed 20-query tree           notebooks/Display_Tree.ipynb (text/plain of cell 2)
  D4     re-rooted ParticleCollection    notebooks/Belief_and_Particle_Filter_Options.ipynb
  D6     push!-hook call count           notebooks/Belief_and_Particle_Filter_Options.ipynb
  D8     Baby reward sanity band         notebooks/Sanity_Checks.ipynb
"""
import json
import os
import re
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json")


def nb_outputs(path):
    nb = json.load(open(path))
    for cell in nb["cells"]:
        if cell["cell_type"] != "code":
            continue
        src = "".join(cell["source"])
        for o in cell.get("outputs", []):
            txt = o.get("text") or o.get("data", {}).get("text/plain") or ""
            yield src, "".join(txt)


def parse_tree(text):
    """Recursive-descent parse of the printed RootNode(...) into nested dicts."""
    i = text.index("POMCP.RootNode")
    pos = [i]

    def expect(tok):
        assert text.startswith(tok, pos[0]), (tok, text[pos[0]:pos[0] + 60])
        pos[0] += len(tok)

    def skip_type():
        # consume an identifier with nested {...} type parameters
        depth = 0
        while True:
            ch = text[pos[0]]
            if ch == "{":
                depth += 1
            elif ch == "}":
                depth -= 1
            elif ch == "(" and depth == 0:
                return
            pos[0] += 1

    def number():
        m = re.match(r"-?\d+(\.\d+)?(e-?\d+)?", text[pos[0]:])
        pos[0] += m.end()
        return float(m.group(0)) if ("." in m.group(0) or "e" in m.group(0)) else int(m.group(0))

    def boolean():
        if text.startswith("true", pos[0]):
            pos[0] += 4
            return True
        expect("false")
        return False

    def belief():
        skip_type()
        expect("(")
        v = number()
        expect(")")
        return v

    def dict_of(parse_val, paired):
        # Dict{K,V}(Pair{..}(k,v),...) or Dict(k=>v,...) or Dict{..}() (empty)
        skip_type()
        expect("(")
        out = []
        while text[pos[0]] != ")":
            if text.startswith("Pair", pos[0]):
                skip_type()
                expect("(")
                k = boolean()
                expect(",")
                v = parse_val()
                expect(")")
            else:
                k = boolean()
                expect("=>")
                v = parse_val()
            out.append((k, v))
            if text[pos[0]] == ",":
                pos[0] += 1
        expect(")")
        return out

    def act_node():
        skip_type()
        expect("(")
        label = boolean()
        expect(",")
        N = number()
        expect(",")
        V = number()
        expect(",")
        kids = dict_of(obs_node, False)
        expect(")")
        return {"a": label, "N": N, "V": V, "obs": [{"o": k, **v} for k, v in kids]}

    def obs_node():
        skip_type()
        expect("(")
        label = boolean()
        expect(",")
        N = number()
        expect(",")
        b = belief()
        expect(",")
        kids = dict_of(act_node, False)
        expect(")")
        return {"N": N, "belief": b, "label": label, "acts": [v for _, v in kids], "act_order": [k for k, _ in kids]}

    skip_type()
    expect("(")
    N = number()
    expect(",")
    b = belief()
    expect(",")
    kids = dict_of(act_node, True)
    return {"N": N, "belief": b, "acts": [v for _, v in kids], "act_order": [k for k, _ in kids]}


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present: golden file is committed, nothing to do")
    g = {"_source": "printed outputs of the notebooks under /root/reference/notebooks (see make_golden.py)"}
    # D3 tree
    for src, out in nb_outputs(os.path.join(REF, "notebooks/Display_Tree.ipynb")):
        if "POMCPTreeVisualizer(POMCP.RootNode" in out:
            g["display_tree"] = parse_tree(out)
            g["display_tree_config"] = {"problem": "BabyPOMDP(-5,-10)", "eps": 0.01, "c": 10.0, "tree_queries": 20,
                                        "rollout": "FeedWhenCrying", "node_belief_updater": "exact"}
            break
    assert "display_tree" in g
    # D1, D4, D6
    pushes = 0
    for src, out in nb_outputs(os.path.join(REF, "notebooks/Belief_and_Particle_Filter_Options.ipynb")):
        m = re.search(r"BoolDistribution\((0\.0240\d+)\)", out)
        if m and "d1_belief" not in g:
            g["d1_belief"] = float(m.group(1))
        m = re.search(r"ParticleCollection\{Bool\}\(Bool\[([a-z,]+)\]\)", out)
        if m and "d4_particles" not in g:
            g["d4_particles"] = [x == "true" for x in m.group(1).split(",")]
            g["d4_tree_queries"] = int(re.search(r"tree_queries=(\d+)", src).group(1)) if "tree_queries" in src else 5
        if "Received state" in out:
            pushes = max(pushes, out.count("Received state"))
    g["d6_push_count_tree_queries_5"] = pushes
    # D2 chain: beliefs that appear in the printed tree
    g["d2_belief_chain"] = {"ff": 0.0240964, "ffff": 0.0298684, "ffffff": 0.0312831, "ffffffff": 0.0316318,
                            "ff_then_ft": 0.525699, "f_t_from_0": 0.470588, "after_feed": 0.0}
    # D8
    rewards = []
    for src, out in nb_outputs(os.path.join(REF, "notebooks/Sanity_Checks.ipynb")):
        m = re.fullmatch(r"\s*(-\d+\.\d+)\s*", out)
        if m:
            rewards.append(float(m.group(1)))
    g["d8_sanity_rewards_in_order"] = rewards
    json.dump(g, open(OUT, "w"), indent=1)
    print("wrote", OUT, "tree root N =", g["display_tree"]["N"], "rewards", rewards)


if __name__ == "__main__":
    main()
