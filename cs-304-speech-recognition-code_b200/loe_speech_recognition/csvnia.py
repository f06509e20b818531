"""The reference's '|'-separated table files (csvnia.py:9-92): truth / prediction tables written by the test
drivers (scripts/project5_test_ndigits_with_sil.py:77-82) and read back by the analysis notebooks.

Format, as the reference writes and parses it: first line = column names; strings are wrapped in double
quotes with embedded quotes doubled, everything else is ``str()``; on reading, a quoted entry becomes a
string, the bare word None becomes ``None``, an all-digit entry an ``int`` and anything else stays a string.
An entry containing '|' or a newline does not survive the round trip (no escaping for those, csvnia.py:41-51).
"""
from __future__ import annotations

import logging
from typing import Dict, List, Union

logger = logging.getLogger(__name__)

SEPARATOR = "|"


class CSV:
    def __init__(self, columns: List[str]) -> None:
        self.columns: List[str] = columns
        self.records: List[List] = []

    def __len__(self) -> int:
        return len(self.records)

    def __str__(self) -> str:
        return f"Columns: {', '.join(self.columns)} Size: {len(self)}"


class CSVWriter(CSV):
    def add_line(self, line: List) -> None:
        self.records.append(line)

    def write(self, path: str) -> None:
        rows = [self.columns] + self.records
        with open(path, "w", encoding="utf-8") as f:
            f.write("".join(self.line_escape(row) + "\n" for row in rows))
        logger.info("Finish writing CSV to %s", path)

    @staticmethod
    def line_escape(line: List) -> str:
        def cell(entry) -> str:
            if isinstance(entry, str):
                return '"' + entry.replace('"', '""') + '"'
            return str(entry)
        return SEPARATOR.join(cell(entry) for entry in line)


class CSVReader(CSV):
    def __init__(self, path: str) -> None:
        with open(path, "r", encoding="utf-8") as f:
            lines = [line.strip() for line in f.readlines()]
        super().__init__([name.replace('"', "") for name in lines[0].split(SEPARATOR)] if lines else [])
        self.records = [self.line_parser(line) for line in lines[1:]]
        self._index: int = -1
        logger.info("Read CSV from %s", path)

    def __iter__(self) -> "CSVReader":
        return self

    def __next__(self) -> Dict[str, Union[str, None, int]]:
        self._index += 1
        if self._index == len(self):
            raise StopIteration
        return dict(zip(self.columns, self.records[self._index]))

    @staticmethod
    def line_parser(line: str) -> List[Union[str, int, None]]:
        parsed: List[Union[str, int, None]] = []
        for entry in line.split(SEPARATOR):
            if entry[0] == '"' and entry[-1] == '"':         # an empty cell raises IndexError, as in the reference
                parsed.append(entry[1:-1].replace('""', '"'))
            elif entry == "None":
                parsed.append(None)
            elif entry.isdigit():
                parsed.append(int(entry))
            else:
                parsed.append(entry)
        return parsed
