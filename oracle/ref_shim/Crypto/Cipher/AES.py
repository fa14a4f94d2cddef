"""`Crypto.Cipher.AES` shim over `cryptography` (OpenSSL).  Only what the reference's
federatedml/secureprotol/jzf_aes.py:33-41 touches: AES.new(key, MODE_ECB).encrypt(block) and the
CTR constructor (seed wrapping, out of scope for the hot path but needed for import)."""
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

MODE_ECB = 1
MODE_CTR = 6


class _Ecb(object):
    def __init__(self, key):
        c = Cipher(algorithms.AES(key), modes.ECB())
        self._enc = c.encryptor()   # one reusable context: per-call cost comparable to pycryptodome
        self._dec = c.decryptor()

    def encrypt(self, data):
        return self._enc.update(data)

    def decrypt(self, data):
        return self._dec.update(data)


class _Ctr(object):
    def __init__(self, key, counter):
        iv = int(counter.get("initial_value", 0)).to_bytes(16, "big")
        c = Cipher(algorithms.AES(key), modes.CTR(iv))
        self._enc = c.encryptor()
        self._dec = c.decryptor()

    def encrypt(self, data):
        return self._enc.update(data)

    def decrypt(self, data):
        return self._dec.update(data)


def new(key, mode, counter=None, **kwargs):
    if mode == MODE_ECB:
        return _Ecb(bytes(key))
    if mode == MODE_CTR:
        return _Ctr(bytes(key), counter or {})
    raise ValueError("shim supports ECB and CTR only")
