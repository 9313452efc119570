"""The four trained Huffman code books (percussive, silence, speech, tonal = table ids 0..3, alphabetical
order).  Data comes from /root/reference/training_data/*_table.pkl via oracle/make_tables.py."""
import json
import os

_here = os.path.dirname(os.path.abspath(__file__))


class HuffTable(object):
    def __init__(self, d):
        self.name = d["name"]
        self.escape = int(d["escape"])
        self.codes = {int(k): str(v) for k, v in d["codes"].items()}    # value -> code string
        self.rev = {v: k for k, v in self.codes.items()}                  # code string -> value
        self.escape_code = self.codes[self.escape]
        self.maxlen = max(len(c) for c in self.codes.values())


def load_tables():
    with open(os.path.join(_here, "huffman_tables.json")) as fh:
        d = json.load(fh)
    return [HuffTable(t) for t in d["tables"]]


TABLES = load_tables()
NO_TABLE = 15
