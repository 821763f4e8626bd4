"""Exception types a caller of the third-party client may catch."""
from retrieval_based_object_detection_b200.store import CollectionNotFound


class ApiException(Exception):
    pass


class UnexpectedResponse(ApiException):
    def __init__(self, status_code=None, reason_phrase="", content=b"", headers=None):
        super().__init__(f"Unexpected Response: {status_code} ({reason_phrase})\nRaw response content:\n{content!r}")
        self.status_code = status_code
        self.reason_phrase = reason_phrase
        self.content = content
        self.headers = headers or {}


class ResponseHandlingException(ApiException):
    def __init__(self, source):
        super().__init__(str(source))
        self.source = source


__all__ = ["ApiException", "UnexpectedResponse", "ResponseHandlingException", "CollectionNotFound"]
