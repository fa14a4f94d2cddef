"""The FATE `Encrypt` plug-in surface (reference: federatedml/secureprotol/encrypt.py:21-79).

Same method names and behaviour: key getters/setters are no-ops in the base class; the list,
table and recursive variants apply encrypt/decrypt element by element."""
from collections.abc import Iterable

import numpy as np


class Encrypt(object):
    def __init__(self):
        self.public_key = None
        self.privacy_key = None

    def generate_key(self, n_length=0):
        pass

    def set_public_key(self, public_key):
        pass

    def get_public_key(self):
        pass

    def set_privacy_key(self, privacy_key):
        pass

    def get_privacy_key(self):
        pass

    def encrypt(self, value):
        pass

    def decrypt(self, value):
        pass

    def encrypt_list(self, values):
        return [self.encrypt(msg) for msg in values]

    def decrypt_list(self, values):
        return [self.decrypt(msg) for msg in values]

    def distribute_decrypt(self, X):
        return X.mapValues(lambda x: self.decrypt(x))

    def distribute_encrypt(self, X):
        return X.mapValues(lambda x: self.encrypt(x))

    def _recursive_func(self, obj, func):
        # encrypt.py:63-73: 1-D arrays element by element, deeper arrays row by row, other
        # iterables rebuilt with their own type, scalars directly.
        if isinstance(obj, np.ndarray):
            if len(obj.shape) == 1:
                return np.reshape([func(val) for val in obj], obj.shape)
            return np.reshape([self._recursive_func(o, func) for o in obj], obj.shape)
        if isinstance(obj, Iterable):
            return type(obj)(self._recursive_func(o, func) if isinstance(o, Iterable) else func(o) for o in obj)
        return func(obj)

    def recursive_encrypt(self, X):
        return self._recursive_func(X, self.encrypt)

    def recursive_decrypt(self, X):
        return self._recursive_func(X, self.decrypt)
