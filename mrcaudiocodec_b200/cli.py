"""Command line front end: WAV -> .pac -> WAV with the GPU codec, in the spirit of the reference's
`python pacfileThem.py in.wav` (pacfileThem.py:1064-1231: encode pass, then decode pass).

    python -m mrcaudiocodec_b200.cli encode in.wav out.pac [--kbps 128] [--independent] [--precision fp64]
                                                            [--block-switching]
    python -m mrcaudiocodec_b200.cli decode in.pac out.wav [--independent]
    python -m mrcaudiocodec_b200.cli roundtrip in.wav            # writes in.pac and in_decoded.wav next to it

WAV handling is host plumbing (stdlib `wave`; 16-bit PCM, two channels: what pcmfile.py:34-66 accepts, minus mono).
--block-switching is the reference's own `__main__` loop (transient detector, one block of look-ahead, eight
128-sample short blocks around transients; the decoder reads the block sizes from the chunk headers).
The .pac files are the reference's format: last block pair non-joint (Q10), so
`decode` needs to be told --independent only for files that were encoded that way (nothing in the file says so).
Decoded WAVs hold (#block pairs) * nMDCTLines frames like the reference's decode loop: the input delayed by nothing
(the first, all-delay block is dropped) and zero-padded to whole blocks plus the flush block."""
import argparse
import os
import sys
import wave

import numpy as np


def read_wav(path):
    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2 or w.getnchannels() != 2 or w.getcomptype() != "NONE":
            raise SystemExit("%s: need 16-bit PCM, two channels" % path)
        sr = w.getframerate()
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").reshape(-1, 2)
    return sr, np.ascontiguousarray(pcm)


def write_wav(path, sr, pcm):
    with wave.open(path, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(int(sr))
        w.writeframes(np.ascontiguousarray(pcm, dtype="<i2").tobytes())


def _codec(sr, args, L=1024, tbps=None, switching=False, n_scale_bits=4, n_mant_size_bits=4):
    from .codec import Codec
    if tbps is None:
        tbps = args.kbps * 1000.0 / sr
    return Codec(sample_rate=sr, n_mdct_lines=L, n_scale_bits=n_scale_bits, n_mant_size_bits=n_mant_size_bits,
                 target_bits_per_sample=tbps, joint=not args.independent,
                 precision=args.precision, device=args.device,
                 block_switching=switching, switch_tables=(L == 1024))


def cmd_encode(args):
    sr, pcm = read_wav(args.input)
    c = _codec(sr, args, switching=bool(args.block_switching))
    blob = c.encode_clips([pcm])[0]
    c.close()
    with open(args.output, "wb") as fh:
        fh.write(blob)
    print("%s: %d frames @ %d Hz -> %s: %d bytes (%.1f kb/s)" %
          (args.input, pcm.shape[0], sr, args.output, len(blob), 8e-3 * len(blob) * sr / max(pcm.shape[0], 1)))


def cmd_decode(args):
    from . import pacfile
    blob = open(args.input, "rb").read()
    h = pacfile.parse_header(blob)
    # the bit rate is not needed to decode; any value builds the same tables.  Field widths come from the header
    # (pacfileThem.py:161-176 reads nScaleBits / nMantSizeBits there: the reference's training files use 3 / 5); the
    # header's band table must equal the one the sample rate and block size give (the library checks it)
    c = _codec(h["sampleRate"], args, L=h["nMDCTLines"], tbps=2.0, n_scale_bits=h["nScaleBits"],
               n_mant_size_bits=h["nMantSizeBits"])
    pcm = c.decode_clips([blob])[0]
    c.close()
    write_wav(args.output, h["sampleRate"], pcm)
    print("%s -> %s: %d frames @ %d Hz" % (args.input, args.output, pcm.shape[0], h["sampleRate"]))


def cmd_roundtrip(args):
    base = os.path.splitext(args.input)[0]
    args.output = base + ".pac"
    cmd_encode(args)
    args.input, args.output = base + ".pac", base + "_decoded.wav"
    cmd_decode(args)


def main(argv=None):
    ap = argparse.ArgumentParser(prog="mrcaudiocodec_b200.cli")
    sub = ap.add_subparsers(dest="cmd", required=True)
    for name, fn, nio in (("encode", cmd_encode, 2), ("decode", cmd_decode, 2), ("roundtrip", cmd_roundtrip, 1)):
        p = sub.add_parser(name)
        p.add_argument("input")
        if nio == 2:
            p.add_argument("output")
        p.add_argument("--kbps", type=float, default=128.0, help="per channel; the reference default is 2.86 bits/sample")
        p.add_argument("--independent", action="store_true", help="independent channels instead of joint M/S")
        p.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
        p.add_argument("--device", type=int, default=0)
        p.add_argument("--block-switching", action="store_true",
                       help="encode with the reference's transient detector and short blocks (joint flow only)")
        p.set_defaults(fn=fn)
    args = ap.parse_args(argv)
    args.fn(args)


if __name__ == "__main__":
    main(sys.argv[1:])
