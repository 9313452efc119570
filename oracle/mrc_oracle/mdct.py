"""MDCT / IMDCT.  Follows /root/reference/mdct.py (MDCTslow :13-50, MDCT :53-96, IMDCT :98-122)."""
import numpy as np


def MDCTslow(data, a, b, isInverse=False):
    """mdct.py:13-50 : O(N^2) definition, 2/N in the forward transform, n0=(b+1)/2."""
    N = a + b
    n0 = (b + 1.0) / 2.0
    n = np.arange(N)
    k = np.arange(N // 2)
    if not isInverse:
        X = np.zeros(N // 2)
        for kk in range(N // 2):
            X[kk] = (2.0 / N) * np.dot(data, np.cos((2.0 * np.pi / N) * np.add(n, n0) * (kk + 1.0 / 2.0)))
        return X
    x = np.zeros(N)
    for nn in range(N):
        x[nn] = np.sum(2.0 * np.asarray(data) * np.cos((2.0 * np.pi / N) * (nn + n0) * (k + 1.0 / 2.0)))
    return x


def MDCT(data, a, b):
    """mdct.py:64-76 : pre-twiddle exp(-j*pi*n/N), N-point complex FFT, post-twiddle
    exp(-j*2*pi*n0*(k+1/2)/N) on bins 0..N/2-1, real part, times 2/N."""
    N = a + b
    n0 = (b + 1.0) / 2.0
    n = np.arange(N)
    pre = np.exp(np.multiply(n, -1j * np.pi / N))
    Y = np.fft.fft(np.multiply(pre, data), N)
    k = np.add(np.arange(N // 2), 1.0 / 2.0)
    post = np.exp(np.multiply(k, -1j * 2.0 * np.pi * n0 / N))
    return (2.0 / N) * np.real(np.multiply(post, Y[0:N // 2]))


def IMDCT(data, a, b):
    """mdct.py:101-118 : X~=[X,-reverse(X)], pre-twiddle exp(j*2*pi*k*n0/N), N-point IFFT, post-twiddle
    exp(j*2*pi*(n+n0)/(2N)), N*real part."""
    N = a + b
    X = np.zeros(N)
    X[0:N // 2] = data
    X[N // 2:] = -1 * np.asarray(data)[::-1]
    n0 = (b + 1) / 2.0
    k = np.arange(N)
    pre = np.exp(np.multiply(k, 1j * 2 * np.pi * n0 / N))
    y = np.fft.ifft(np.multiply(pre, X), N)
    post = np.exp(np.multiply(np.add(k, n0), (1j * 2 * np.pi / (2.0 * N))))
    return N * np.real(np.multiply(y, post))
