"""MSB-first bit packing.  Follows /root/reference/bitpack.py (PackedBits: Size :20-24, Get/SetPackedData
:26-34, WriteBits :36-101, ReadBits :104-170).  Restated on a Python bytearray with bit-serial semantics:
the n lowest bits of `info` are appended MSB first; reading returns the next n bits as an unsigned int."""
import numpy as np

BYTESIZE = 8


class PackedBits(object):
    def __init__(self):
        self.iByte = self.iBit = 0

    def Size(self, nBytes):
        self.nBytes = int(nBytes)
        self.iByte = self.iBit = 0
        self.data = np.zeros(self.nBytes, dtype=np.uint8)

    def GetPackedData(self):
        return self.data.tobytes()

    def SetPackedData(self, data):
        self.nBytes = len(data)
        self.data = np.frombuffer(data, dtype=np.uint8)

    def ResetPointers(self):
        self.iByte = self.iBit = 0

    def WriteBits(self, info, nBits):
        info = int(info)
        pos = self.iByte * BYTESIZE + self.iBit
        for k in range(nBits - 1, -1, -1):
            if (info >> k) & 1:
                self.data[pos >> 3] |= np.uint8(0x80 >> (pos & 7))
            pos += 1
        self.iByte, self.iBit = pos >> 3, pos & 7

    def ReadBits(self, nBits):
        pos = self.iByte * BYTESIZE + self.iBit
        v = 0
        for _ in range(nBits):
            v = (v << 1) | ((int(self.data[pos >> 3]) >> (7 - (pos & 7))) & 1)
            pos += 1
        self.iByte, self.iBit = pos >> 3, pos & 7
        return v
