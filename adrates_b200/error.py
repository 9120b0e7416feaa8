"""Library error type (mirrors cavour/utils/error.py: LibError(message) with `_message`)."""


class LibError(Exception):
    def __init__(self, message: str, code: int = 0):
        super().__init__(message)
        self._message = message
        self.code = code          # native CAV_E_* code when the error comes from the CUDA library, else 0

    def _print(self):
        print("LibError:", self._message)
