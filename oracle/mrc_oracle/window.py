"""Windows.  Follows /root/reference/window.py (SineWindow :10-25, HanningWindow :28-45, KBDWindow :49-102,
TransitionWindow :104-121).  The windows are data independent, so the coefficient vectors are built once per
length and cached; the arithmetic that builds them is the reference's (same numpy calls, same order), which
makes the cached vector bit-identical to what the reference recomputes on every call."""
import numpy as np

_cache = {}


def sine_coeffs(N):
    """window.py:19-23 : sin(pi/N * (n+0.5))."""
    key = ("sine", N)
    if key not in _cache:
        n = np.add(np.linspace(0, N - 1, N), 0.5)
        _cache[key] = np.sin(np.multiply(np.pi / N, n))
    return _cache[key]


def hann_coeffs(N):
    """window.py:36-42 : 0.5 + (-0.5)*cos((2*pi/N) * (n+0.5))."""
    key = ("hann", N)
    if key not in _cache:
        n = np.add(np.linspace(0, N - 1, N), 0.5)
        c = np.cos(np.multiply((2.0 * np.pi) / N, n))
        _cache[key] = np.add(0.5, np.multiply(-0.5, c))
    return _cache[key]


def kbd_coeffs(N, alpha=4.0):
    """window.py:57-99.  M=N/2; v[j]=I0(pi*alpha*sqrt(1-((j-M/2)/(M/2))^2))/I0(pi*alpha), j=0..M;
    top half  w[n]   = sqrt( (tril(ones) . v^2[0:M])[n] / sum(v^2) )   (np.dot => BLAS summation order)
    bottom    w[M+n] = sqrt( (triu(ones) . v^2[1:M+1])[n] / sum(v^2) )."""
    key = ("kbd", N, alpha)
    if key not in _cache:
        M = N / 2.0
        j = np.linspace(0, M, int(M + 1))
        denom0 = np.i0(np.pi * alpha)
        j = np.subtract(j, M / 2.0)
        j = np.square(np.divide(j, M / 2.0))
        rad = np.sqrt(np.subtract(1.0, j))
        v = np.divide(np.i0(np.multiply(np.pi * alpha, rad)), denom0)
        v2 = np.square(v)
        top_in = v2[0:np.size(v2) - 1]
        bot_in = v2[1:np.size(v2)]
        h = int(N / 2.0)
        denom = np.sum(v2)
        top = np.sqrt(np.divide(np.dot(np.tril(np.ones((h, h))), top_in), denom))
        bot = np.sqrt(np.divide(np.dot(np.triu(np.ones((h, h))), bot_in), denom))
        _cache[key] = np.concatenate((top, bot))
    return _cache[key]


def transition_coeffs(a, b):
    """window.py:112-119 : first a samples use the left half of KBD(2a), last b the right half of KBD(2b)."""
    key = ("trans", a, b)
    if key not in _cache:
        _cache[key] = np.append(kbd_coeffs(2 * a)[:a], kbd_coeffs(2 * b)[b:])
    return _cache[key]


def SineWindow(x):
    return np.multiply(x, sine_coeffs(np.size(x)))


def HanningWindow(x):
    return np.multiply(x, hann_coeffs(np.size(x)))


def KBDWindow(x, alpha=4.0):
    return np.multiply(x, kbd_coeffs(np.size(x), alpha))


def TransitionWindow(x, a, b):
    """window.py:104-121.  The reference windows [x[:a],0..0] and [0..0,x[a:]] separately and splices; the
    zero halves never reach the output, so this is x * transition_coeffs(a,b) element for element."""
    return np.multiply(x, transition_coeffs(a, b))
