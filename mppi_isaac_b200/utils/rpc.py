"""world <-> planner RPC: the zerorpc request/reply protocol restated on pyzmq + msgpack.

The reference exposes the planner to the simulated world with ``zerorpc.Server(planner).bind("tcp://0.0.0.0:4242")``
(``examples/panda/planner.py:46-48``) and drives it with ``zerorpc.Client().connect("tcp://127.0.0.1:4242")`` /
``planner.compute_action_tensor(dof_bytes, root_bytes)`` (``examples/panda/world.py:21-22,35-50``).  zerorpc is not a
dependency of this repository; the two classes below speak its wire protocol (version 3) so that either side can be
swapped for the original:

* transport: the server is a ZeroMQ ROUTER, the client a DEALER; a message is ``[identity..., b"", payload]`` on the ROUTER
  side and ``[b"", payload]`` on the DEALER side;
* payload: ``msgpack([header, name, args])`` with ``header = {"message_id": id, "v": 3}`` plus ``"response_to": request id``
  on everything that belongs to the channel a request opened;
* a call is the event ``name = method, args = [positional arguments]``; the answer is ``"OK"`` with ``args = [result]`` or
  ``"ERR"`` with ``args = [exception name, message, traceback]``;
* ``"_zpc_hb"`` heart-beats are answered in kind (a plan takes milliseconds, far below zerorpc's 10 s lost-remote limit),
  ``"_zpc_more"`` flow-control events are accepted and ignored (a reply is a single event).

PARITY NOTE: restated from the published protocol; zerorpc itself is not installed in this image, so the classes are
tested against each other and against hand-built frames (``tests/test_host_logic.py``), not against zerorpc.
"""
import threading
import traceback
import uuid

import msgpack
import zmq

PROTOCOL_VERSION = 3


def _pack(header, name, args) -> bytes:
    return msgpack.packb([header, name, list(args)], use_bin_type=True)


def _unpack(blob: bytes):
    header, name, args = msgpack.unpackb(blob, raw=False)
    return header, name, args


def _new_id() -> str:
    return uuid.uuid4().hex


class RpcServer:
    """Serves the public methods of ``target`` (the reference passes the ``MPPIisaacPlanner``).  Single-threaded like the
    reference's server: calls are executed in arrival order on the thread that runs :meth:`run`."""

    def __init__(self, target, context: zmq.Context = None):
        self.target = target
        self._ctx = context or zmq.Context.instance()
        self._sock = self._ctx.socket(zmq.ROUTER)
        self._sock.setsockopt(zmq.LINGER, 0)
        self._stop = threading.Event()
        self.calls = 0

    def bind(self, endpoint: str):
        self._sock.bind(endpoint)
        return self

    @property
    def last_endpoint(self) -> str:
        return self._sock.getsockopt(zmq.LAST_ENDPOINT).decode()

    def _reply(self, identity, request_id, name, args):
        header = {"message_id": _new_id(), "v": PROTOCOL_VERSION, "response_to": request_id}
        self._sock.send_multipart(list(identity) + [b"", _pack(header, name, args)])

    def handle_one(self, timeout_ms: int = 100) -> bool:
        """Process at most one event; returns False on timeout."""
        if not self._sock.poll(timeout_ms):
            return False
        parts = self._sock.recv_multipart()
        identity, blob = parts[:-2], parts[-1]
        try:
            header, name, args = _unpack(blob)
            if not isinstance(header, dict) or not isinstance(name, str) or not isinstance(args, (list, tuple)):
                return True                  # a malformed frame must not take the (single-threaded) server down: drop it
            request_id = header.get("response_to") or header["message_id"]
        except Exception:
            return True                      # not a zerorpc event: drop it
        if name == "_zpc_hb":
            self._reply(identity, request_id, "_zpc_hb", [])
            return True
        if name == "_zpc_more":
            return True
        if name.startswith("_") or not callable(getattr(self.target, name, None)):
            self._reply(identity, request_id, "ERR", ["NameError", f"no such method: {name}", ""])
            return True
        try:
            result = getattr(self.target, name)(*args)
            self.calls += 1
            self._reply(identity, request_id, "OK", [result])
        except Exception as e:               # the error travels to the caller, the server keeps serving
            self._reply(identity, request_id, "ERR", [type(e).__name__, str(e), traceback.format_exc()])
        return True

    def run(self):
        while not self._stop.is_set():
            self.handle_one(100)
        self._sock.close()

    def stop(self):
        self._stop.set()


class RemoteError(RuntimeError):
    pass


class RpcClient:
    """``client.method(*args)`` performs a blocking remote call (what ``zerorpc.Client`` gives the reference's world)."""

    def __init__(self, endpoint: str = None, timeout_s: float = 30.0, context: zmq.Context = None):
        self._ctx = context or zmq.Context.instance()
        self._sock = self._ctx.socket(zmq.DEALER)
        self._sock.setsockopt(zmq.LINGER, 0)
        self._timeout_ms = int(timeout_s * 1000)
        if endpoint:
            self.connect(endpoint)

    def connect(self, endpoint: str):
        self._sock.connect(endpoint)
        return self

    def close(self):
        self._sock.close()

    def _call(self, name, args):
        msg_id = _new_id()
        self._sock.send_multipart([b"", _pack({"message_id": msg_id, "v": PROTOCOL_VERSION}, name, args)])
        while True:
            if not self._sock.poll(self._timeout_ms):
                raise TimeoutError(f"no answer to {name}() within {self._timeout_ms / 1000:.1f} s")
            header, rname, rargs = _unpack(self._sock.recv_multipart()[-1])
            if header.get("response_to") != msg_id or rname in ("_zpc_hb", "_zpc_more"):
                continue                      # heart-beat or an answer to an abandoned call
            if rname == "OK":
                return rargs[0]
            if rname == "ERR":
                raise RemoteError(f"{rargs[0]}: {rargs[1]}")
            raise RemoteError(f"unexpected event {rname!r}")

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return lambda *args: self._call(name, args)
