"""Filter banks in PyWavelets orientation ``(dec_lo, dec_hi, rec_lo, rec_hi)``.

PyWavelets itself is not a dependency of this package (the reference pulls it in, ``requirements.txt:12``); the
decomposition low-pass filters below are the published Daubechies / symlet / coiflet coefficients, and the other
three filters of an orthogonal bank follow from the quadrature-mirror relations.  ``IUWT`` is the bank the
reference defines itself (dictionaries/undecimated.py:20-25), with the band-pass ``g`` in the low-pass slot as
there.  Any other bank can be passed explicitly as ``filter_bank=[dec_lo, dec_hi, rec_lo, rec_hi]``.
"""
import numpy as np

_DEC_LO = {
    "db1": [0.7071067811865476, 0.7071067811865476],
    "db2": [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025],
    "db3": [0.035226291882100656, -0.08544127388224149, -0.13501102001039084, 0.4598775021193313,
            0.8068915093133388, 0.3326705529509569],
    "db4": [-0.010597401784997278, 0.032883011666982945, 0.030841381835986965, -0.18703481171888114,
            -0.02798376941698385, 0.6308807679295904, 0.7148465705525415, 0.23037781330885523],
    "coif1": [-0.01565572813546454, -0.0727326195128539, 0.38486484686420286, 0.8525720202122554,
              0.3378976624578092, -0.0727326195128539],
    "coif2": [-0.0007205494453645122, -0.0018232088707029932, 0.0056114348193944995, 0.023680171946334084,
              -0.0594344186464569, -0.0764885990783064, 0.41700518442169254, 0.8127236354455423,
              0.3861100668211622, -0.06737255472196302, -0.04146493678175915, 0.016387336463522112],
    "sym4": [-0.07576571478927333, -0.02963552764599851, 0.49761866763201545, 0.8037387518059161,
             0.29785779560527736, -0.09921954357684722, -0.012603967262037833, 0.0322231006040427],
}
_ALIASES = {"haar": "db1", "sym2": "db2", "sym3": "db3"}


def _biorthogonal_2_2():
    a, b, c, r = 0.1767766952966369, 0.3535533905932738, 1.0606601717798214, 0.7071067811865476
    return ([0.0, -a, b, c, b, -a], [0.0, b, -r, b, 0.0, 0.0], [0.0, b, r, b, 0.0, 0.0], [0.0, a, b, -c, b, a])


def _iuwt():
    h = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
    g = [-1 / 16, -4 / 16, 10 / 16, -4 / 16, -1 / 16]
    delta = [0.0, 0.0, 1.0, 0.0, 0.0]
    # odd-length banks are padded with one trailing zero so the filter length is even
    return tuple(f + [0.0] for f in (g, h, delta, delta))


def available(kind="all"):
    names = sorted(list(_DEC_LO) + list(_ALIASES) + ["bior2.2"])
    return names + ["IUWT"] if kind == "all" else names


def filter_bank(name=None, filter_bank=None):
    """four float64 arrays of equal, even length"""
    if filter_bank is not None:
        bank = [np.asarray(f, dtype=np.float64) for f in filter_bank]
        if len(bank) != 4 or len({b.shape[0] for b in bank}) != 1:
            raise ValueError("filter_bank must be four filters of equal length")
        if bank[0].shape[0] % 2:
            bank = [np.concatenate([b, [0.0]]) for b in bank]
        return tuple(bank)
    if name == "IUWT":
        return tuple(np.asarray(f, dtype=np.float64) for f in _iuwt())
    if name == "bior2.2":
        return tuple(np.asarray(f, dtype=np.float64) for f in _biorthogonal_2_2())
    key = _ALIASES.get(name, name)
    if key not in _DEC_LO:
        raise NotImplementedError("The wavelet has not been implemented: %r (available: %s)" % (name, available()))
    dec_lo = np.asarray(_DEC_LO[key], dtype=np.float64)
    taps = np.arange(dec_lo.shape[0])
    rec_lo = dec_lo[::-1].copy()
    dec_hi = np.where(taps % 2 == 0, -1.0, 1.0) * rec_lo
    rec_hi = np.where(taps % 2 == 0, 1.0, -1.0) * dec_lo
    return dec_lo, dec_hi, rec_lo, rec_hi


def swt_max_level(length):
    """how many times ``length`` halves exactly"""
    level = 0
    while length > 0 and length % 2 == 0:
        length //= 2
        level += 1
    return level


def dwt_max_level(length, filter_len):
    if filter_len < 2 or length < filter_len - 1:
        return 0
    return int(np.floor(np.log2(length / (filter_len - 1.0))))
