"""
make_tables.py -- dump the reference's four trained Huffman code books to JSON (run once, in the build
container, where /root/reference exists).  TEST INFRASTRUCTURE / data preparation only.

Source: /root/reference/training_data/{percussive,silence,speech,tonal}_table.pkl (value -> (code
string, length), escape value), cross-checked against *_table.revpkl (code -> value, escape code) and
*_tree.pkl (HuffmanNode tree, walked here) -- the three artefacts the reference reads at
codecThem.py:137-147, pacfileThem.py:170-171,243-256.  Canonical table order is alphabetical
(SURVEY.md Appendix C, Q3): 0 percussive, 1 silence, 2 speech, 3 tonal.

Writes the same JSON to oracle/mrc_oracle/huffman_tables.json (the oracle's copy) and
mrcaudiocodec_b200/huffman_tables.json (the product's copy; constant data, like the .pkl files are
constant data for the reference).
"""
import json
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

NAMES = ["percussive", "silence", "speech", "tonal"]


def walk(node, prefix, out, HuffmanNode):
    """node is a (payload, freq) tuple; payload is a HuffmanNode or a leaf value."""
    payload = node[0]
    if isinstance(payload, HuffmanNode):
        walk(payload.left, prefix + "0", out, HuffmanNode)
        walk(payload.right, prefix + "1", out, HuffmanNode)
    else:
        out[prefix] = int(payload)


def main():
    mods = ref_shim.load()
    HuffmanNode = mods["huffman"].HuffmanNode
    td = os.path.join(ref_shim.REF_DIR, "training_data")
    tables = []
    for name in NAMES:
        with open(os.path.join(td, name + "_table.pkl"), "rb") as fh:
            table, esc = pickle.load(fh, encoding="latin1")
        with open(os.path.join(td, name + "_table.revpkl"), "rb") as fh:
            rev, esc_code = pickle.load(fh, encoding="latin1")
        with open(os.path.join(td, name + "_tree.pkl"), "rb") as fh:
            root, esc2 = pickle.load(fh, encoding="latin1")
        codes = {int(v): str(c[0]) for v, c in table.items()}
        for v, c in table.items():
            assert len(c[0]) == c[1], (name, v, c)
        assert {str(k): int(v) for k, v in rev.items()} == {c: v for v, c in codes.items()}, name
        assert str(esc_code) == codes[int(esc)], name
        assert str(esc2[0]) == codes[int(esc)] and int(esc2[1]) == len(esc2[0]), name
        tree_codes = {}
        walk(root, "", tree_codes, HuffmanNode)
        assert tree_codes == {c: v for v, c in codes.items()}, name
        kraft = sum(2.0 ** -len(c) for c in codes.values())
        assert abs(kraft - 1.0) < 1e-12, (name, kraft)
        tables.append({"name": name, "escape": int(esc),
                       "codes": {str(v): codes[v] for v in sorted(codes)}})
        print(name, "escape", esc, "n", len(codes), "maxlen", max(len(c) for c in codes.values()))
    blob = json.dumps({"order": NAMES, "tables": tables}, indent=1, sort_keys=True)
    for dst in (os.path.join(HERE, "mrc_oracle", "huffman_tables.json"),
                os.path.join(HERE, "..", "mrcaudiocodec_b200", "huffman_tables.json")):
        with open(dst, "w") as fh:
            fh.write(blob + "\n")
        print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    main()
