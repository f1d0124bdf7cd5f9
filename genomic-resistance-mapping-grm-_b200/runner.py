"""Host utilities of the GUI's process plumbing, for this path only -- function forms of
``run_bash_command`` (src/util.py:123-154) and ``display_process_output`` (src/util.py:196-243).

The reference writes the command into ``<temp>/tmp.sh`` (which deletes itself), starts it through ``wsl -e`` and pumps
stdout / stderr line by line into a Tk text box, tagging stderr.  Here the script runs under ``bash`` directly (there is
no Windows host) and the sink is a callable; everything else -- the temp script, the self-removal, the tags, "exit status
0 = success" (src/app.py:3406) -- is the same.
"""
from __future__ import annotations

import os
from queue import Empty, Queue
from subprocess import PIPE, Popen
from threading import Thread
from time import sleep
from typing import Callable, Optional

from .kover_cmd import to_linux_path


class Tag:
    NORMAL = "normal"
    ERROR = "error"
    SUCCESS = "success"


def run_bash_command(command: str, temp_path: Optional[str] = None) -> Optional[Popen]:
    file_name = "tmp.sh"
    if temp_path:
        os.makedirs(temp_path, exist_ok=True)
    temp_file = os.path.join(temp_path, file_name) if temp_path else file_name
    with open(temp_file, "wb") as bash_file:
        bash_file.write(f'#!/bin/bash\n{command}\nrm "$0"'.encode("UTF-8").replace(b"\r\n", b"\n"))
    os.chmod(temp_file, 0o755)
    return Popen(["bash", to_linux_path(temp_file).strip('"')], stdout=PIPE, stderr=PIPE, text=True)


def _enqueue_output(stream, queue: Queue, tag: str) -> None:
    def pump():
        for line in iter(stream.readline, ""):
            queue.put((tag, line))
        stream.close()
    Thread(target=pump, daemon=True).start()


def display_process_output(process: Popen, output_target: Optional[Callable[[str, str], None]] = None,
                           refresh_timeout: int = 0, message_buffer: int = 1) -> int:
    """Pump the child's output to ``output_target(text, tag)`` (default: print, stderr lines prefixed with their tag)
    until it exits; returns the exit status."""
    messages: Queue = Queue()
    _enqueue_output(process.stdout, messages, Tag.NORMAL)
    _enqueue_output(process.stderr, messages, Tag.ERROR)

    def emit(text, tag):
        if output_target:
            output_target(text, tag)
        else:
            print(text if tag == Tag.NORMAL else f"{tag}: {text}", end="")

    while process.poll() is None or not messages.empty():
        try:
            tag, message = messages.get(timeout=0.1)
            emit(message, tag)
        except Empty:
            pass
        sleep(refresh_timeout / 1000)
    sleep(0.05)
    while not messages.empty():
        tag, message = messages.get()
        emit(message, tag)
    return process.returncode
