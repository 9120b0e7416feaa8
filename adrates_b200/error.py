"""Library error type (mirrors cavour/utils/error.py: LibError(message) with `_message`)."""


class LibError(Exception):
    def __init__(self, message: str):
        super().__init__(message)
        self._message = message

    def _print(self):
        print("LibError:", self._message)
