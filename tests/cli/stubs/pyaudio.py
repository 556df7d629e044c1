"""Test infrastructure: stand-in for the `pyaudio` module the reference's support/cli.py imports at the top
(reference support/cli.py:12).  With `-o file` the CLI never opens an audio device (cli.py:150-157)."""
paInt16 = 8


class PyAudio:
    def get_default_output_device_info(self):
        raise OSError("no audio device (pyaudio stub)")

    def open(self, **kw):
        raise OSError("no audio device (pyaudio stub)")

    def terminate(self):
        pass
