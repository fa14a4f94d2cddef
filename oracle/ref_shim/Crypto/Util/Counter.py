"""`Crypto.Util.Counter` shim: the reference only calls Counter.new(128, initial_value=0)
(federatedml/secureprotol/jzf_aes.py:32)."""


def new(nbits, initial_value=0, **kwargs):
    return {"nbits": nbits, "initial_value": initial_value}
