// Parser.h -- tokenizer of rtigo3's two text formats (system and scene description).
// Grammar restated from apps/rtigo3/src/Parser.cpp:72-226: blanks/tabs separate tokens, '#' starts a
// comment that runs to the end of the line, a token is a VALUE when it starts like a number and holds
// only number characters, getNextLine() returns the rest of the line (paths with blanks).
#pragma once
#include <string>

enum ParserTokenType { PTT_UNKNOWN, PTT_ID, PTT_VAL, PTT_STRING, PTT_EOL, PTT_EOF };

class Parser
{
public:
  bool load(std::string const& filename);
  void setSource(std::string const& text) { m_source = text; m_index = 0; m_line = 1; }
  ParserTokenType getNextToken(std::string& token);
  ParserTokenType getNextLine(std::string& token);
  std::string::size_type getSize() const { return m_source.size(); }
  std::string::size_type getIndex() const { return m_index; }
  unsigned int getLine() const { return m_line; }

private:
  std::string m_source;
  std::string::size_type m_index = 0;
  unsigned int m_line = 1;
};
