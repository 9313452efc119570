"""
mrc_oracle -- CPU restatement (numpy, Python 3) of the laser55/mrcAudioCodec encode/decode hot path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package, and only as the checker / the timed CPU baseline.  The product package
(mrcaudiocodec_b200) never imports it and has no CPU fallback.

Pinning: this restatement is pinned (a) by the reference's own known-answer tests -- TDAC 12-sample vector
(mdct.py:131,174-179), fast-vs-slow MDCT/IMDCT at N=1024 (mdct.py:184-210), bit-pack vector 0x3AB7
(bitpack.py:183-196) -- and (b) by golden vectors under tests/golden/ produced by running the UNMODIFIED
reference sources through oracle/ref_shim.py (oracle/make_golden.py is the committed generating script):
.pac bytes and every per-block intermediate (overall scales, ms_switch, bit allocations, scale factors,
mantissas, table ids, reservoir) for joint and independent-channel encodes, and decoded PCM.
Each function cites the reference file:line it follows.  Quirks Q1-Q12 of SURVEY.md Appendix C are
reproduced on purpose.
"""
from . import window, mdct, psychoac, bitalloc, quantize, ms_stereo, bitpack, tables, codec, pacfile, pcm, driver, huffman_train, transient  # noqa: F401
